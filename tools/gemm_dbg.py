"""Diagnostic: time the tcgen05 GEMM on the backbone's big shapes with parts of the epilogue disabled
(HMDPOSE_GEMM_DBG bit 0: no useful global stores, bit 1: no math/stores, bit 2: no TMEM loads)."""
import ctypes
import os
import subprocess
import sys

import numpy as np

SHAPES = [(262144, 96, 16), (65536, 144, 24), (65536, 24, 144), (262144, 16, 32), (16384, 240, 40), (16384, 40, 240), (4096, 480, 80)]


def child():
    from hmd_ego_pose_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(0)
    out = []
    for (M, N, K) in SHAPES:
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = np.zeros(N, np.float32)
        D = np.zeros((M, N), np.float32)
        ms = ctypes.c_float(0)
        rc = lib.hmdpose_test_gemm(0, 1, 1, M, N, K, A.ctypes.data, W.ctypes.data, bias.ctypes.data, None, M, None, 1,
                                   D.ctypes.data, ctypes.byref(ms))
        out.append(f"{ms.value * 1000:7.1f}" if rc == 0 else f"rc={rc}")
    print(os.environ.get("HMDPOSE_GEMM_DBG", "0"), " ".join(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        print("dbg  " + " ".join(f"{m}x{n}x{k}" for m, n, k in SHAPES))
        for dbg in sys.argv[2:] or ["0", "1", "2", "6"]:
            env = dict(os.environ, HMDPOSE_GEMM_DBG=dbg)
            subprocess.run([sys.executable, __file__, "child"], env=env, check=False)
