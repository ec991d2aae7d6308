"""Staged GPU diagnostics (run on the B200 box): each section runs in its own process so that a
trapped kernel cannot poison the following sections.  Writes gpurun_out/check_<section>.txt.

    python tools/gpu_check.py all
"""
import os
import subprocess
import sys
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
GOLD = os.path.join(ROOT, "tests", "golden")


def relerr(a, b):
    import numpy as np
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-12))


def sec_gemm(log):
    import ctypes
    import numpy as np
    from hmd_ego_pose_b200 import _native
    lib = _native.load()
    rng = np.random.default_rng(0)

    def run(impl, prec, M, N, K, gate=False, res=False, act=0, rpi=0):
        A = rng.standard_normal((M, K)).astype(np.float32)
        W = (rng.standard_normal((N, K)) / np.sqrt(K)).astype(np.float32)
        bias = rng.standard_normal(N).astype(np.float32) * 0.1
        rpi_ = rpi if rpi else M
        nimg = (M + rpi_ - 1) // rpi_
        g = (rng.random((nimg, K)).astype(np.float32) + 0.25) if gate else None
        R = rng.standard_normal((M, N)).astype(np.float32) if res else None
        D = np.zeros((M, N), np.float32)
        ms = ctypes.c_float(0)
        if prec == 1:  # what the kernel sees after fp16 rounding of the operands
            A16, W16 = A.astype(np.float16).astype(np.float64), W.astype(np.float16).astype(np.float64)
        else:
            A16, W16 = A.astype(np.float64), W.astype(np.float64)
        if gate:
            rows = np.arange(M) // rpi_
            A16 = A16 * g[rows].astype(np.float64)
            if prec == 1 and impl >= 1:
                A16 = A16.astype(np.float16).astype(np.float64)
        ref = A16 @ W16.T + bias
        if act == 1:
            ref = ref / (1 + np.exp(-ref))
        elif act == 2:
            ref = 1 / (1 + np.exp(-ref))
        if res:
            ref = ref + (R.astype(np.float16).astype(np.float64) if prec == 1 else R)
        rc = lib.hmdpose_test_gemm(0, impl, prec, M, N, K, A.ctypes.data, W.ctypes.data, bias.ctypes.data,
                                   g.ctypes.data if gate else None, rpi_, R.ctypes.data if res else None, act,
                                   D.ctypes.data, ctypes.byref(ms))
        if rc != 0:
            return f"rc={rc} {lib.hmdpose_last_error(None)}"
        return f"err={relerr(D, ref):.2e} ms={ms.value:.4f}"

    shapes = [(262144, 96, 16), (65536, 144, 24), (65536, 24, 144), (16384, 240, 40), (128, 64, 64), (256, 96, 16), (300, 144, 24), (1000, 40, 144), (4096, 1152, 192), (640, 320, 1152),
              (64, 64, 64), (4, 64, 64), (2048, 16, 32), (5000, 240, 40), (16384, 96, 16), (777, 112, 672),
              (1364, 64, 320), (262144, 96, 16), (65536, 144, 24)]
    for (M, N, K) in shapes:
        log(f"M={M} N={N} K={K}: simt32 {run(0, 0, M, N, K)} | tc16v2 {run(1, 1, M, N, K)} | tc16v1 {run(2, 1, M, N, K)}")
    for (M, N, K, rpi) in [(1024, 24, 96, 256), (640, 320, 1152, 64), (1000, 80, 480, 100), (4096, 16, 32, 4096)]:
        log(f"gated M={M} N={N} K={K} rpi={rpi}: simt32 {run(0, 0, M, N, K, True, True, 0, rpi)} | "
            f"tc16v2 {run(1, 1, M, N, K, True, True, 0, rpi)} | tc16v1 {run(2, 1, M, N, K, True, True, 0, rpi)}")
    for act in (1, 2):
        log(f"act={act} M=512 N=64 K=64: simt32 {run(0, 0, 512, 64, 64, act=act)} | tc16 {run(1, 1, 512, 64, 64, act=act)}")


def _weights():
    from oracle import synth_weights as sw
    return sw.synthetic_weights(0, 256, bn_stats=sw.load_bn_stats(os.path.join(GOLD, "bn_stats_seed0.npz")))


def _stages(log, precision, S=256, B=2, env=None):
    import numpy as np
    import torch
    for k, v in (env or {}).items():
        os.environ[k] = v
    os.environ["HMDPOSE_KEEP_ALL"] = "1"
    from hmd_ego_pose_b200 import HmdPoseSession
    from oracle import net_ref
    sd = _weights()
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1234))
    probe = net_ref.forward_probe(sd, x)
    sess = HmdPoseSession(sd, image_size=S, max_batch=B, precision=precision, use_graph=False)
    t0 = time.time()
    outs = sess.forward_raw(x.cuda())
    torch.cuda.synchronize()
    log(f"forward_raw ok in {time.time() - t0:.3f}s launches={sess.last_launch_count}")

    def cmp(name, ref_nchw):
        got = sess.debug_read(name)
        ref = ref_nchw.permute(0, 2, 3, 1).contiguous().numpy().ravel()
        if got.size != ref.size:
            log(f"  {name}: SIZE MISMATCH {got.size} vs {ref.size}")
            return
        log(f"  {name:28s} relerr={relerr(got, ref):.3e}  nan={int(np.isnan(got).sum())}")
    cmp("stem", probe["stem_blocks"][0])
    for i in range(16):
        cmp(f"blk{i}", probe["stem_blocks"][i + 1])
    for c in range(3):
        for l in range(5):
            cmp(f"cell{c}.p{l + 3}", probe["cells"][c][l])
    for n, got, ref in zip(("regression", "classification", "rotation", "translation_raw", "hand"), outs, probe["out"]):
        log(f"  out {n:20s} relerr={relerr(got.cpu().numpy(), ref.numpy()):.3e}")
    return sess, sd, x, probe


def sec_parity(log):
    _stages(log, "parity")


def sec_fast_simt(log):
    _stages(log, "fast", env={"HMDPOSE_FORCE_SIMT": "1"})


def sec_fast_tc(log):
    _stages(log, "fast")


def sec_parity512(log):
    _stages(log, "parity", S=512, B=1)


def sec_post(log):
    import numpy as np
    import torch
    from hmd_ego_pose_b200 import HmdPoseSession
    from oracle import net_ref, postprocess_ref as pp
    sd = _weights()
    B, S = 4, 256
    x = torch.randn(B, 3, S, S, generator=torch.Generator().manual_seed(1234))
    feats, reg, cls, rot, tr, hand = [t.numpy() if not isinstance(t, tuple) else t for t in net_ref.forward(sd, x)]
    cam = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))
    cam[1] = [687.7084, 688.8967, 435.8758, 242.4822, 1000, 1]
    sess = HmdPoseSession(sd, image_size=S, max_batch=B, precision="parity")
    ref = pp.detect(reg, cls, rot, tr, hand, cam, S)
    got = sess.postprocess_host(reg, cls, rot, tr, hand, cam)
    for b in range(B):
        r = ref[b]
        log(f"image {b}: count={int(r['count'])} idx_equal={np.array_equal(got['anchor_idx'][b], r['anchor_idx'])} "
            f"labels_equal={np.array_equal(got['labels'][b], r['labels'])} "
            f"scores_equal={np.array_equal(got['scores'][b], r['scores'])} "
            f"box_err={np.abs(got['boxes'][b] - r['boxes']).max():.2e} "
            f"trans_err={np.abs(got['translation'][b] - r['translation']).max():.2e} "
            f"rot_equal={np.array_equal(got['rotation'][b], r['rotation'])} hand_equal={np.array_equal(got['hand'][b], r['hand'])}")
    # identical pre-NMS inputs -> bit-exact everything
    a, t = pp.anchors_for_shape((S, S))
    boxes = pp.decode_boxes(a, reg, S, S)
    trans = pp.decode_translation(t, tr, cam)
    got2 = sess.filter_boxes_host(boxes, cls, rot, trans, hand)
    for b in range(B):
        r = pp.filter_detections(boxes[b], cls[b], rot[b], trans[b], hand[b])
        ok = all(np.array_equal(got2[k][b], r[k]) for k in ("boxes", "scores", "labels", "rotation", "translation", "hand", "anchor_idx"))
        log(f"filter_boxes image {b}: bit-exact={ok}")
    for b in range(B):
        r = pp.csharp_best(reg[b], cls[b], rot[b], tr[b], cam[b], S)
        g = sess.best_from_raw_host(reg[b], cls[b], rot[b], tr[b], cam[b])
        log(f"best image {b}: ref={np.round(r, 4).tolist()} err={np.abs(g - r).max():.2e}")
    # end to end through the net in parity mode
    det = sess.detect_host(x.numpy(), cam)
    for b in range(B):
        r = ref[b]
        log(f"e2e parity image {b}: idx_equal={np.array_equal(det['anchor_idx'][b], r['anchor_idx'])} "
            f"n={int(r['count'])} box_err={np.abs(det['boxes'][b] - r['boxes']).max():.2e} "
            f"rot_err={np.abs(det['rotation'][b] - r['rotation']).max():.2e} trans_err={np.abs(det['translation'][b] - r['translation']).max():.2e}")
    log(f"launches={sess.last_launch_count} gpu_ms={sess.last_gpu_ms:.3f}")


def sec_speed(log):
    import numpy as np
    import torch
    from hmd_ego_pose_b200 import HmdPoseSession
    sd = _weights()
    for precision in ("parity", "fast"):
        for (B, S, mb) in [(1, 256, 0), (16, 256, 0), (16, 256, 8), (64, 256, 16), (16, 512, 4), (64, 512, 8)]:
            try:
                sess = HmdPoseSession(sd, image_size=S, max_batch=B, precision=precision, micro_batch=mb)
                x = torch.randn(B, 3, S, S, device="cuda")
                cam = torch.tensor([[480., 480., 128., 128., 1000., 1.]], device="cuda").repeat(B, 1)
                for _ in range(3):
                    sess.detect(x, cam)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 10
                e0.record()
                for _ in range(n):
                    sess.detect(x, cam)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / n
                log(f"{precision} B={B} S={S} mb={mb}: {ms:.3f} ms/step  {B / ms * 1000:.0f} frames/s launches={sess.last_launch_count}")
                sess.close()
            except Exception as e:
                log(f"{precision} B={B} S={S} mb={mb}: FAILED {e}")


def sec_steps(log):
    import numpy as np
    import torch
    from hmd_ego_pose_b200 import HmdPoseSession
    sd = _weights()
    B = int(os.environ.get("STEPS_B", "16"))
    sess = HmdPoseSession(sd, image_size=256, max_batch=B, precision="fast")
    x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(1234)).numpy()
    cam = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))
    sess.detect_host(x, cam)
    prof = sess.profile_steps(B, mode=1, reps=10)
    tot = sum(p[2] for p in prof)
    log(f"B={B} total {tot:.3f} ms over {len(prof)} launches (un-graphed, event per launch)")
    for name, kern, ms, by, fl in prof:
        log(f"{name:34s} {kern:20s} {ms * 1e3:8.1f} us {by / 1e6:9.2f} MB {by / max(ms, 1e-9) / 1e6:8.1f} GB/s {fl / max(ms, 1e-9) / 1e9:7.2f} TF/s")
    tl = sess.debug_read("__s3_timeline")
    names = ["entry", "prologue done", "pdl_wait done", "weights issued", "weights landed", "first tile landed",
             "first accf (grp 0)", "second accf (grp 1)", "last tile MMA start", "last tile accf", "roles done", "exit"]
    log("sepconv3 CTA 0 timeline (us since entry, last launch = header): " + ", ".join(f"{n}={v:.2f}" for n, v in zip(names, tl)))
    cn = ["step start", "fence+sync", "phase A done", "phase B done", "MMA done", "step end"]
    for st in range(2):
        log(f"sepconv chain CTA 0 step {st} (us since chain start): " + ", ".join(f"{n}={tl[16 + st * 8 + i]:.2f}" for i, n in enumerate(cn)))


def sec_insitu(log):
    """In-situ cost of every step: time of the graph of steps[0..k] minus that of steps[0..k-1] (PDL and L2 state as in
    the real plan), next to the isolated event-per-launch time."""
    import numpy as np
    import torch
    from hmd_ego_pose_b200 import HmdPoseSession
    sd = _weights()
    B = int(os.environ.get("STEPS_B", "16"))
    sess = HmdPoseSession(sd, image_size=256, max_batch=B, precision="fast")
    x = torch.randn(B, 3, 256, 256, generator=torch.Generator().manual_seed(1234)).numpy()
    cam = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))
    sess.detect_host(x, cam)
    iso = sess.profile_steps(B, mode=1, reps=10)
    ins = sess.profile_steps(B, mode=1 | 0x100, reps=20)
    # marginal cost of every launch with 8 steps in flight (8 streams replaying the same prefix graph)
    fl8 = sess.profile_steps(B, mode=1 | 0x200, reps=10) if os.environ.get("INFLIGHT8", "1") != "0" else [(0, 0, 0.0, 0, 0)] * len(ins)
    log(f"B={B} in-situ total {sum(p[2] for p in ins):.3f} ms, isolated total {sum(p[2] for p in iso):.3f} ms, "
        f"8 in flight {sum(p[2] for p in fl8):.3f} ms per step, {len(ins)} launches")
    cat, cat8 = {}, {}
    for (name, kern, ms, by, fl), (_, _, ms_iso, _, _), (_, _, ms8, _, _) in zip(ins, iso, fl8):
        log(f"{name:34s} {kern:20s} in-situ {ms * 1e3:7.1f} us  isolated {ms_iso * 1e3:7.1f} us  in-flight-8 {ms8 * 1e3:7.1f} us  {by / 1e6:8.2f} MB {by / max(ms, 1e-9) / 1e6:8.1f} GB/s")
        key = name.split(".")[-1] if name.startswith("blk") else name.split(".")[0]
        cat[key] = cat.get(key, 0.0) + ms
        cat8[key] = cat8.get(key, 0.0) + ms8
    log("by category (in-situ us): " + ", ".join(f"{k}={v * 1e3:.0f}" for k, v in sorted(cat.items(), key=lambda kv: -kv[1])))
    log("by category (8 in flight, us per step): " + ", ".join(f"{k}={v * 1e3:.0f}" for k, v in sorted(cat8.items(), key=lambda kv: -kv[1])))


SECTIONS = {"insitu": sec_insitu, "gemm": sec_gemm, "parity": sec_parity, "fast_simt": sec_fast_simt, "fast_tc": sec_fast_tc,
            "parity512": sec_parity512, "post": sec_post, "speed": sec_speed, "steps": sec_steps}


def main():
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "all":
        for name in SECTIONS:
            t0 = time.time()
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], cwd=ROOT, timeout=900)
            print(f"[{name}] rc={r.returncode} {time.time() - t0:.1f}s", flush=True)
        return
    path = os.path.join(OUT, f"check_{which}.txt")
    with open(path, "w") as f:
        def log(s):
            print(s, flush=True)
            f.write(s + "\n")
            f.flush()
        try:
            SECTIONS[which](log)
        except Exception:
            log("EXCEPTION\n" + traceback.format_exc())
            sys.exit(1)


if __name__ == "__main__":
    main()
