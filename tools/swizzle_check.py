"""Brute-force check of the shared-memory bank behaviour of the DMMA fragment loads in starneig_b200/csrc/dgemm_tma.cuh.

TMA writes tiles with CU_TENSOR_MAP_SWIZZLE_128B: rows of 16 doubles (128 bytes), the 16-byte chunk index of an element XORed
with (row & 7). Lane 4g+t of a DMMA.8x8x4 holds A(x = base + g, k = K(s, t)); an 8-byte LDS of a warp is served in two phases of
16 lanes, conflict-free iff the 16 lanes hit 16 distinct 8-byte words modulo 128 bytes. With the natural order k = 4s + t both
operand layouts have 2-way conflicts; with the permuted k sets of tma_kperm() both are conflict-free.   usage: swizzle_check.py"""
KSETS = [[0, 3, 12, 15], [1, 2, 13, 14], [4, 7, 8, 11], [5, 6, 9, 10]]          # tma_kperm(s, t)


def addr_mn(x, k):      # MN-major tile: boxes of 16 x-values (inner) x 16 k (rows), side by side (tma_tile_offset<false>)
    box, xi = divmod(x, 16)
    return box * 256 + k * 16 + ((((xi >> 1) ^ (k & 7)) << 1) | (xi & 1))


def addr_k(x, k):       # K-major tile: 16 k (inner) x BX rows (tma_tile_offset<true>)
    return x * 16 + ((((k >> 1) ^ (x & 7)) << 1) | (k & 1))


def worst(addr, ksets):
    w = 0
    for base in range(0, 128, 8):
        for ks in ksets:
            for half in (0, 1):
                words = [addr(base + g, ks[t]) % 16 for g in range(4 * half, 4 * half + 4) for t in range(4)]
                w = max(w, max(words.count(v) for v in set(words)))
    return w


if __name__ == "__main__":
    natural = [[4 * s + t for t in range(4)] for s in range(4)]
    assert sorted(k for ks in KSETS for k in ks) == list(range(16))
    print("natural k order : MN-major %d-way, K-major %d-way" % (worst(addr_mn, natural), worst(addr_k, natural)))
    print("permuted k sets : MN-major %d-way, K-major %d-way" % (worst(addr_mn, KSETS), worst(addr_k, KSETS)))
    assert worst(addr_mn, KSETS) == 1 and worst(addr_k, KSETS) == 1
