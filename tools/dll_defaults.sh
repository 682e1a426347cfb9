# StartRT .. WaitRT with the reference's ParamsRT defaults (YulioRT.h:37-50: size 1536, spp 256, depth 10) on the generated room scene
# usage: bash tools/dll_defaults.sh <gpus>
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from tests import dae_scene
dae_scene.write_scene("/tmp/dll", "room", tex_size=256)
PY
YULIO_RT_CFG=gpus=$1 yulio_raytracer_b200/lib/rt_test /tmp/dll/room.dae 1536 256 10 2>&1 | grep -v "^state" | tail -3
ls -la /tmp/dll | tail -2
