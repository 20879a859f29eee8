"""N > 1 on real GPUs: torchrun with 2 ranks, the distributed step (owner-computes all-to-all and the
all-gather flavour) equals the single-GPU step bit for bit.  Skipped on boxes with one GPU."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("region_cap", [None, "256"])
def test_two_ranks_equal_one(region_cap):
    """region_cap = 256: the receive regions of the in-library exchange start far too small, so the first steps overflow and
    the ranks grow, re-map (cudaIpc) and repeat the step together -- same bits in the end."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641" if region_cap is None else "29642", os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ)
    if region_cap:
        env["CLSN_DIST_REGION_CAP"] = region_cap
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("dist ok") == 2


def test_cpp_host_mirror_two_gpus_in_one_process(tmp_path):
    """The multi-GPU step through the reference-shaped C++ API: one host thread per GPU, each with its own mesh copy and
    its own CollisionSolver3d(device), joined by enableMultiGPU (peer access inside the process instead of cudaIpc).
    Every rank's result equals the single-GPU run bit for bit."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from collision_b200 import scenes
    from parity_util import same_bits
    from test_gpu_host_cpp import write_scene
    host = os.path.join(ROOT, "collision_b200", "host")
    subprocess.check_call(["make", "-s", "-C", host])
    exe = os.path.join(host, "host_check")
    for sc in (scenes.layered_cloth(4, 24, seed=99), scenes.ball_plane(gap=2e-4)):
        inp, out1, out2 = str(tmp_path / "scene.bin"), str(tmp_path / "one.bin"), str(tmp_path / "two.bin")
        write_scene(sc, inp)
        subprocess.check_call([exe, inp, out1, "3", "1", "nozones"])
        subprocess.check_call([exe, inp, out2, "3", "2"], timeout=300)
        one = np.fromfile(out1, dtype=np.float64)
        for r in range(2):
            assert same_bits(np.fromfile(out2 + f".{r}", dtype=np.float64), one), (sc.name, r)
