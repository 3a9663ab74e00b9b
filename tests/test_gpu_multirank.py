"""N-rank = 1-rank on real GPUs: the dry core with the grid tracer and the moist model on 2 (and, when the box
has them, 4) ranks -- peer-memory transposes fused into the Legendre / FFT epilogues, scalar all-reduces, tracer halo exchange --
must reproduce the single-GPU run on every grid and spectral field, `divs` included, to 1e-10 (relative to the field maximum) after
40 steps from a cold start (by then the Held-Suarez forcing has produced a real divergent circulation; right after a cold start `divs`
is round-off of a non-divergent flow and has no meaningful relative error).  Skipped on a box with one GPU; the host-side layouts are
covered without GPUs by tests/test_multirank_cpu.py (gloo, world size 2)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("ranks,transport", [(2, "p2p"), (2, "nccl"), (4, "p2p")])
def test_n_ranks_reproduce_one_rank(lib_built, ranks, transport):
    if _gpus() < ranks:
        pytest.skip(f"needs {ranks} GPUs")
    env = dict(os.environ)
    if transport == "nccl":
        env["ISCA_B200_NO_P2P"] = "1"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multigpu_check.py"), "T42", "20", "40", "6"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("MULTIGPU_JSON ")]
    assert line, r.stdout[-2000:]
    rep = json.loads(line[-1][len("MULTIGPU_JSON "):])
    assert rep["ranks"] == ranks
    for k, v in rep["dry"].items():
        assert v < 1e-10, ("dry", k, v, rep)
    for k, v in rep["moist"].items():
        assert v < 1e-10, ("moist", k, v, rep)
