"""Host-side model of the BVH8 node's leaf-triangle mask and of the slot permutation (csrc/bvh.cuh: node_test, perm8, tri_index;
csrc/bvh_build.cu: k_collapse). The kernels themselves are checked on the GPU against the oracle (tests/test_gpu_parity.py,
test_gpu_atsize.py); this file pins the bit arithmetic they rely on, so that a change of the node format that breaks one of the
identities fails here, without a GPU.

    triMask bit 8k + s   = slot s is a leaf child with more than k triangles (k < 3)
    storage order        = ascending bit number: triangle of bit b is triBase + popc(triMask & ((1 << b) - 1))
    triangle bits of a test = (hitSlots & ~imask) * 0x010101 & triMask
    traversal order      = slot s is visited at position s ^ octinv, highest position first
"""
import random


def perm8(x, o):          # csrc/bvh.cuh: perm8
    if o & 1:
        x = ((x & 0x55) << 1) | ((x >> 1) & 0x55)
    if o & 2:
        x = ((x & 0x33) << 2) | ((x >> 2) & 0x33)
    if o & 4:
        x = ((x & 0x0F) << 4) | (x >> 4)
    return x


def popc(x):
    return bin(x).count("1")


def make_node(rng):
    """Random slot assignment: kind[s] in {'empty', 'inner', 'leaf'}, cnt[s] triangles for a leaf."""
    kind, cnt = [], []
    for _ in range(8):
        k = rng.choice(["empty", "inner", "leaf", "leaf"])
        kind.append(k)
        cnt.append(rng.randint(1, 3) if k == "leaf" else 0)
    imask = sum(1 << s for s in range(8) if kind[s] == "inner")
    tri_mask = 0
    for s in range(8):
        for k in range(cnt[s]):
            tri_mask |= 1 << (8 * k + s)
    return kind, cnt, imask, tri_mask


def test_perm8_moves_slot_s_to_position_s_xor_octant():
    for o in range(8):
        seen = set()
        for x in range(256):
            want = sum(((x >> s) & 1) << (s ^ o) for s in range(8))
            assert perm8(x, o) == want
            seen.add(want)
        assert len(seen) == 256            # a permutation of the masks
        for x in range(256):
            assert perm8(perm8(x, o), o) == x      # XOR with a constant is an involution


def test_highest_position_first_visits_the_slots_in_octant_order():
    rng = random.Random(5)
    for _ in range(200):
        o, hits = rng.randrange(8), rng.randrange(256)
        ordered, visited = perm8(hits, o), []
        while ordered:
            bit = ordered.bit_length() - 1
            ordered &= ~(1 << bit)
            visited.append(bit ^ o)         # trace_stream: slot = (bit - 24) ^ octinv
        assert sorted(visited) == [s for s in range(8) if (hits >> s) & 1]
        assert [s ^ o for s in visited] == sorted((s ^ o for s in visited), reverse=True)


def test_tri_mask_numbering_is_a_bijection_onto_the_nodes_triangles():
    rng = random.Random(7)
    for _ in range(500):
        kind, cnt, imask, tri_mask = make_node(rng)
        n = sum(cnt)
        assert popc(tri_mask) == n and tri_mask < (1 << 24)
        # the builder writes triangle k of slot s at triBase + popc(triMask & below(8k + s)); the traversal reads the same index
        index = {}
        for s in range(8):
            for k in range(cnt[s]):
                b = 8 * k + s
                index[(s, k)] = popc(tri_mask & ((1 << b) - 1))
        assert sorted(index.values()) == list(range(n))
        # plane-major order: every first triangle comes before every second one
        firsts = [index[(s, 0)] for s in range(8) if cnt[s] > 0]
        seconds = [index[(s, 1)] for s in range(8) if cnt[s] > 1]
        assert not seconds or max(firsts) < min(seconds)
        assert imask & (tri_mask & 0xFF) == 0          # a slot is inner or leaf, never both


def test_triangle_bits_of_a_node_test():
    rng = random.Random(11)
    for _ in range(500):
        kind, cnt, imask, tri_mask = make_node(rng)
        hits = rng.randrange(256)           # may include empty slots (a degenerate node lets an inverted box through) and inner slots
        got = (((hits & ~imask) & 0xFF) * 0x010101) & tri_mask
        want = 0
        for s in range(8):
            if (hits >> s) & 1 and kind[s] == "leaf":
                for k in range(cnt[s]):
                    want |= 1 << (8 * k + s)
        assert got == want
        inner = hits & imask
        assert all(kind[s] == "inner" for s in range(8) if (inner >> s) & 1)


def test_sign_bit_mask_assembly():
    """node_test shifts the sign of (tmax - tmin) of slot 7, 6, ... 0 into a mask: after eight funnel shifts slot s sits at bit s."""
    rng = random.Random(13)
    for _ in range(200):
        miss = [rng.random() < 0.5 for _ in range(8)]
        acc = 0
        for s in range(7, -1, -1):
            sign = 0x80000000 if miss[s] else 0
            acc = ((acc << 1) | (sign >> 31)) & 0xFFFFFFFF      # __funnelshift_l(sign word, acc, 1)
        hit8 = ~acc & 0xFF
        assert hit8 == sum((0 if miss[s] else 1) << s for s in range(8))
