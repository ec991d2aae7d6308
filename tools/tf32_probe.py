"""Accuracy probe of the parity-mode GEMMs (FFMA impl 0 vs 3xTF32 tcgen05 impl 3) against fp64:
max / rms relative error and the mean SIGNED error (a truncating accumulator shows up as a bias towards zero).
python tools/tf32_probe.py [net]   -- `net` also prints the end-to-end head errors of a parity session."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hmd_ego_pose_b200 import _native
lib = _native.load()

def run(impl, A, W, b, act=0):
    M, K = A.shape; N = W.shape[0]
    D = np.zeros((M, N), np.float32); ms = ctypes.c_float()
    rc = lib.hmdpose_test_gemm(0, impl, 0, M, N, K, A.ctypes.data, W.ctypes.data, b.ctypes.data, None, M, None, act,
                               D.ctypes.data, ctypes.byref(ms))
    assert rc == 0, lib.hmdpose_last_error(None)
    return D, ms.value

rng = np.random.default_rng(0)
for M, N, K, pos in [(4096, 96, 16, 0), (4096, 144, 24, 0), (4096, 40, 240, 0), (2048, 112, 672, 0), (2048, 192, 1152, 0),
                     (2048, 320, 1152, 0), (2048, 64, 320, 0), (2048, 192, 1152, 1), (2048, 64, 576, 1)]:
    A = rng.standard_normal((M, K)).astype(np.float32)
    W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
    if pos:  # all-positive operands: no cancellation, the accumulator grows monotonically
        A, W = np.abs(A), np.abs(W)
    b = np.zeros(N, np.float32)
    ref = A.astype(np.float64) @ W.astype(np.float64).T
    for impl in (0, 3):
        D, ms = run(impl, A, W, b)
        e = D.astype(np.float64) - ref
        print(f"M={M} N={N} K={K} pos={pos} impl={impl}: max {np.abs(e).max() / np.abs(ref).max():.2e}  "
              f"rms {np.sqrt((e ** 2).mean()) / np.sqrt((ref ** 2).mean()):.2e}  "
              f"signed {(e * np.sign(ref)).mean() / np.abs(ref).mean():+.2e}  ({ms * 1e3:.1f} us)")

if len(sys.argv) > 1 and sys.argv[1] == "net":
    import torch
    from hmd_ego_pose_b200 import HmdPoseSession
    from oracle import net_ref, synth_weights as sw
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sd = sw.synthetic_weights(0, 256, bn_stats=sw.load_bn_stats(os.path.join(ROOT, "tests", "golden", "bn_stats_seed0.npz")))
    x = torch.randn(4, 3, 256, 256, generator=torch.Generator().manual_seed(7))
    ref = [t.numpy() for t in net_ref.forward(sd, x)[1:]]
    s = HmdPoseSession(sd, image_size=256, max_batch=4, precision="parity")
    got = s.raw_host(x.numpy())
    tag = "FORCE_SIMT" if os.environ.get("HMDPOSE_FORCE_SIMT") else "tf32"
    for name, g, r in zip(("regression", "classification", "rotation", "translation_raw", "hand"), got, ref):
        print(f"net[{tag}] {name}: relerr {np.abs(g - r).max() / np.abs(r).max():.3e}  max abs {np.abs(g - r).max():.3e}")
