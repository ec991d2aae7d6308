"""Concurrency stress of the BENCHED configuration: batch 16, fast mode, five handles on five CUDA
streams driven by five host threads -- exactly what bench.py's `value` and `e2e` legs do -- for hundreds of steps, and
every result must be BITWISE equal to the same batch run alone on a fresh handle (the kernels are deterministic).
Run once with the spinning host wait and once with HMDPOSE_BLOCKING_SYNC=1 (the switch bench.py flips when
ranks x callers exceed the host cores, i.e. the 8-GPU run that died with CUDA 719 in round 1).

A protocol race, a cross-stream TMEM deadlock (the device watchdog would trap and hmdpose_last_error would name the
barrier) or any scheduling-dependent result shows up here as a mismatch or an exception."""
import os
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

B, S = 16, 256
KEYS = ("boxes", "scores", "labels", "rotation", "translation", "hand", "anchor_idx")
CAM = np.tile(np.array([[480, 480, 128, 128, 1000, 1]], np.float32), (B, 1))


def _pool(n):
    g = torch.Generator().manual_seed(4321)
    return [torch.randn(B, 3, S, S, generator=g).numpy() for _ in range(n)]


def _make(sd, n, blocking):
    from hmd_ego_pose_b200 import HmdPoseSession
    old = os.environ.get("HMDPOSE_BLOCKING_SYNC")
    if blocking:
        os.environ["HMDPOSE_BLOCKING_SYNC"] = "1"
    else:
        os.environ.pop("HMDPOSE_BLOCKING_SYNC", None)
    try:   # the switch is read when a handle is created
        return [HmdPoseSession(sd, image_size=S, max_batch=B, precision="fast") for _ in range(n)]
    finally:
        if old is None:
            os.environ.pop("HMDPOSE_BLOCKING_SYNC", None)
        else:
            os.environ["HMDPOSE_BLOCKING_SYNC"] = old


@pytest.fixture(scope="module")
def serial_results(synth_sd):
    """every pool batch run alone, one call at a time, on its own handle"""
    pool = _pool(4)
    s = _make(synth_sd, 1, False)[0]
    ref = [s.detect_host(x, CAM) for x in pool]
    again = [s.detect_host(x, CAM) for x in pool]
    for a, b in zip(ref, again):
        for k in KEYS:
            assert np.array_equal(a[k], b[k]), k     # run-to-run determinism is the premise of this test
    assert sum(int((r["labels"] >= 0).sum()) for r in ref) > 0
    s.close()
    return pool, ref


@pytest.mark.parametrize("blocking", [False, True], ids=["spin", "blocking_sync"])
def test_five_host_threads_five_handles_bitwise(synth_sd, serial_results, blocking):
    """bench.py's e2e leg: hmdpose_run_detect with host buffers from 5 threads, 300 steps each."""
    pool, ref = serial_results
    n_thr, steps = 5, 300
    sessions = _make(synth_sd, n_thr, blocking)
    errors = []

    def worker(k):
        try:
            for i in range(steps):
                j = (i + k) % len(pool)
                got = sessions[k].detect_host(pool[j], CAM)
                for key in KEYS:
                    if not np.array_equal(got[key], ref[j][key]):
                        errors.append((k, i, key))
                        return
        except Exception as e:   # noqa: BLE001  (a trapped kernel surfaces here with the watchdog record)
            errors.append((k, "exception", repr(e)))

    ths = [threading.Thread(target=worker, args=(k,)) for k in range(n_thr)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    for s in sessions:
        s.close()
    assert not errors, errors[:5]


@pytest.mark.parametrize("blocking", [False, True], ids=["spin", "blocking_sync"])
def test_five_streams_round_robin_device_api_bitwise(synth_sd, serial_results, blocking):
    """bench.py's `value` leg: hmdpose_run_detect_device on 5 torch streams, steps issued round-robin, 300 steps."""
    pool, ref = serial_results
    n, steps = 5, 300
    dev = torch.device("cuda", 0)
    sessions = _make(synth_sd, n, blocking)
    streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
    d_pool = [torch.from_numpy(x).to(dev) for x in pool]
    cam = torch.from_numpy(CAM).to(dev)
    torch.cuda.synchronize(dev)
    outs = []
    for i in range(steps):
        k = i % n
        with torch.cuda.stream(streams[k]):
            outs.append((i % len(pool), sessions[k].detect(d_pool[i % len(pool)], cam)))
    torch.cuda.synchronize(dev)
    bad = []
    for i, (j, o) in enumerate(outs):
        for key, t in zip(KEYS, o):
            if not np.array_equal(t.cpu().numpy(), ref[j][key]):
                bad.append((i, key))
                break
    for s in sessions:
        s.close()
    assert not bad, bad[:5]


def test_mixed_host_and_device_calls_on_one_handle_are_ordered(synth_sd, serial_results):
    """a handle used through the device API on a torch stream and through the host API (the handle's own
    stream) without any host synchronisation in between must still give each call its own results."""
    pool, ref = serial_results
    dev = torch.device("cuda", 0)
    s = _make(synth_sd, 1, False)[0]
    side = torch.cuda.Stream(device=dev)
    d_pool = [torch.from_numpy(x).to(dev) for x in pool]
    cam = torch.from_numpy(CAM).to(dev)
    torch.cuda.synchronize(dev)
    for i in range(20):
        j = i % len(pool)
        with torch.cuda.stream(side):
            o = s.detect(d_pool[j], cam)                       # asynchronous, on `side`
        h = s.detect_host(pool[(j + 1) % len(pool)], CAM)      # handle stream, starts immediately
        side.synchronize()
        for key, t in zip(KEYS, o):
            assert np.array_equal(t.cpu().numpy(), ref[j][key]), (i, key, "device call")
        for key in KEYS:
            assert np.array_equal(h[key], ref[(j + 1) % len(pool)][key]), (i, key, "host call")
    assert s.last_gpu_ms > 0
    s.close()
