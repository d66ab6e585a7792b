"""N>1 host logic on CPU: two gloo ranks (no GPU needed)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_rank_gloo_halo_logic():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29733", os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "DIST CPU OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_two_rank_gloo_dropin_host_logic():
    """Multi-rank drop-in layer without a device: PCDKSP's ownership offsets (exscan), the
    distributed sub-matrix extraction, the row-partitioned PCDAssembler and the sub-field BC index
    mapping with its split offset (SubfieldBC.h:138-140) against the serial objects."""
    env = dict(os.environ, OMP_NUM_THREADS="2")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29735", os.path.join(ROOT, "tests", "dist_dropin_worker.py"),
           "--host-only"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "DROPIN HOST OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]
