"""Host logic of the launch plans of the fused / split-K kernels (hmd_ego_pose_b200/csrc/common.cuh): compiled with nvcc
as a host-only program and run on the CPU (no GPU, no CUDA call)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


@pytest.mark.skipif(not (os.path.exists(NVCC) or shutil.which("nvcc")), reason="nvcc not available")
def test_fused_kernel_plans_respect_sm100a_budgets(tmp_path):
    exe = str(tmp_path / "plan_check")
    nvcc = NVCC if os.path.exists(NVCC) else shutil.which("nvcc")
    cmd = [nvcc, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "hmd_ego_pose_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
           "-o", exe, os.path.join(ROOT, "tests", "native", "plan_check.cu")]
    subprocess.run(cmd, check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=300)
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=60)
    print(r.stdout)
    assert r.returncode == 0, r.stdout
    last = r.stdout.strip().splitlines()[-1]
    assert "failures 0" in last
    # blocks 6-15 at 256x256 (fused MBConv), blocks 1-5 at both sizes (expand + depthwise), and the latency plans of the split-K kernel
    counts = dict(zip(("mbconv", "expdw", "projk"), [int(t.rstrip(",")) for t in last.replace("plans:", "").split() if t.rstrip(",").isdigit()][:3]))
    assert counts["mbconv"] >= 10 and counts["expdw"] == 10 and counts["projk"] >= 30
