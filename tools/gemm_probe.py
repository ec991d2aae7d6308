"""Single big pointwise GEMM through hmdpose_test_gemm (for ncu): python tools/gemm_probe.py [impl] [M N K]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import _native
lib = _native.load()
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 1
M, N, K = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (262144, 96, 16)
act = int(sys.argv[5]) if len(sys.argv) > 5 else 1
rng = np.random.default_rng(0)
A = rng.standard_normal((M, K)).astype(np.float32); W = rng.standard_normal((N, K)).astype(np.float32) / np.sqrt(K)
b = np.zeros(N, np.float32); D = np.zeros((M, N), np.float32); ms = ctypes.c_float()
for _ in range(3):
    rc = lib.hmdpose_test_gemm(0, impl, 1, M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, None, M, None, act, D.ctypes.data, ctypes.byref(ms))
    print(rc, ms.value, "ms", (M * K * 2 + M * N * 2) / ms.value / 1e6, "GB/s")
