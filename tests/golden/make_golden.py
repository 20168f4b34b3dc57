"""Generates tests/golden/*.npz from the REFERENCE's own Hessenberg sources.

Must be run in the authoring container (needs /root/reference): `make -C oracle` builds
oracle/_ref/libstarneig_ref.so from /root/reference/src/{hessenberg,common}/*.c against the sequential
StarPU stand-in (oracle/ref_shim), and this script runs starneig_SEP_SM_Hessenberg[_expert] from that
library on reference-test-driver inputs (LCG fullpos / partial generators, test/common/init.c,
test/misc/partial_hessenberg.c) and stores inputs' seeds + outputs. The fixtures travel to the GPU box,
/root/reference does not.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle.oracle import Oracle, Reference  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# (name, n, begin, end, tile_size, panel_width, generator, seed)
CASES = [
    ("full_n47_t16_p8", 47, 0, 47, 16, 8, "fullpos", 2019),
    ("full_n130_default", 130, 0, 130, -1, -1, "fullpos", 7),
    ("full_n200_t48_p35", 200, 0, 200, 48, 35, "full", 11),
    ("partial_n88_t24_p16", 88, 22, 66, 24, 16, "partial", 2019),
    ("partial_n150_t40_p45", 150, 37, 112, 40, 45, "partial", 3),
]


def main():
    ora, ref = Oracle(), Reference()
    ref.set_threads(1)
    for name, n, begin, end, tile, pw, gen, seed in CASES:
        if gen == "fullpos":
            A, Q, ld = ora.fullpos(n, seed)
        elif gen == "full":
            A, Q, ld = ora.full(n, seed)
        else:
            A, Q, ld = ora.partial(n, begin, end, seed)
        A0 = A.copy(order="F")
        ret = ref.hessenberg_expert(n, A, ld, Q, ld, begin, end, tile, pw)
        assert ret == 0
        np.savez_compressed(os.path.join(HERE, name + ".npz"), n=n, begin=begin, end=end, tile_size=tile,
                            panel_width=pw, generator=gen, seed=seed, ld=ld, A0=A0[:n], H=A[:n], Q=Q[:n])
        print(name, "residual %.1f u" % ora.residual_u(n, Q, ld, A, ld, A0, ld), "orth %.1f u" % ora.orthogonality_u(n, Q, ld))
    # LCG known answers (test/common/common.c:48-59): first values for seed 2019
    ora.prand_init(2019)
    seq = np.array([ora.prand() for _ in range(16)], dtype=np.int64)
    np.savez(os.path.join(HERE, "prand_seed2019.npz"), seq=seq)
    print("prand", seq[:4])


if __name__ == "__main__":
    main()
