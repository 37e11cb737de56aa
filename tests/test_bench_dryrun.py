"""Control-flow check of bench.py without a GPU: `bench.main()` runs on the CPU with the oracle operator backend patched
in and the CUDA timing primitives faked, at a tiny geometry.  The numbers mean nothing; what is checked is that every
code path of the default run, of the eager run and of the video benchmark executes and prints a line that carries every
key of the bench contract (a NameError in this script at round end would cost the round's measurement)."""
import contextlib
import io
import json
import os
import sys
import time

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.cpu_backend import oracle_ops  # noqa: E402


@pytest.fixture(autouse=True)
def _restore_policy():
    """bench.main() sets the process-wide arithmetic policy; put the library default back for the tests that follow"""
    yield
    from univs_b200.precision import set_precision
    set_precision("fp32")


class _Event:
    def __init__(self, enable_timing=False):
        self.t = None

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

    def synchronize(self):
        pass


class _Stream:
    cuda_stream = 0

    def synchronize(self):
        pass

    def wait_stream(self, other):
        pass


@contextlib.contextmanager
def _fake_cuda(monkeypatch):
    monkeypatch.setenv("UNIVS_BENCH_DEVICE", "cpu")
    monkeypatch.setenv("UNIVS_CPU_THREADS", "4")
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setitem(bench.WORKLOADS, "dry", ("tiny", 2, 64, 96, 8))
    monkeypatch.setitem(bench.PROMPT_WORKLOADS, "dry_sot", ("tiny", 2, 64, 96, 8, "sot", 1))
    monkeypatch.setitem(bench.PROMPT_WORKLOADS, "dry_grounding", ("tiny", 2, 64, 96, 8, "grounding", 0))
    with oracle_ops():
        yield


def _run(monkeypatch, argv):
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    buf = io.StringIO()
    with _fake_cuda(monkeypatch), contextlib.redirect_stdout(buf):
        bench.main()
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1, buf.getvalue()
    return json.loads(lines[0])


CONTRACT = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline")


def test_default_run_control_flow(monkeypatch):
    # --no-graph: CUDA-graph capture is the one part that cannot be imitated on a CPU (runtime.GraphedClip)
    line = _run(monkeypatch, ["--workload", "dry", "--steps", "1", "--warmup", "1", "--no-graph", "--precision", "fp32"])
    for k in CONTRACT:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 3       # W >= 3 is enforced
    assert line["value"] > 0 and line["unit"] == "frames/s" and line["higher_is_better"] is True
    from univs_b200 import switches
    assert line["config"]["workload"].startswith("dry:") and line["config"]["switches"] == switches.active()
    assert line["config"]["execution"] == "eager" and "l2" in line["config"]
    e2e = line["e2e"]
    assert e2e["value"] > 0 and e2e["h2d_bytes_per_step"] == 2 * 3 * 64 * 96 and e2e["d2h_bytes_per_step"] > 0
    cpu = line["cpu_baseline"]
    assert cpu["kind"] == "port" and cpu["cores"] >= 1 and cpu["value"] > 0 and "sample" in cpu
    assert line["named_kernel_rooflines"] == {}         # no CUDA brackets on this backend: absent, not wrong


def test_video_benchmark_control_flow(monkeypatch):
    line = _run(monkeypatch, ["--workload", "dry", "--steps", "1", "--video-frames", "3", "--no-graph", "--precision", "fp32"])
    assert line["metric"].startswith("video frames/sec") and line["value"] > 0
    assert line["reference_schedule"]["value"] > 0 and line["rle_results"]["value"] > 0


def test_plain_cpu_run_fails_loudly(monkeypatch):
    """without the oracle patch the same invocation must die in the first operator: there is no CPU path"""
    from univs_b200._cabi import UnivsB200Error
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", "dry", "--steps", "1", "--no-graph"])
    monkeypatch.setenv("UNIVS_BENCH_DEVICE", "cpu")
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setitem(bench.WORKLOADS, "dry", ("tiny", 2, 64, 96, 8))
    monkeypatch.setitem(bench.PROMPT_WORKLOADS, "dry_sot", ("tiny", 2, 64, 96, 8, "sot", 1))
    monkeypatch.setitem(bench.PROMPT_WORKLOADS, "dry_grounding", ("tiny", 2, 64, 96, 8, "grounding", 0))
    with pytest.raises(UnivsB200Error):
        bench.main()


def _sharded_worker(rank, world, port, out_dir):
    """one rank of `torchrun ... bench.py --gpus 2` on the CPU: gloo, oracle operators, faked CUDA timers"""
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port), UNIVS_BENCH_DEVICE="cpu", UNIVS_CPU_THREADS="2")
    torch.set_num_threads(2)
    torch.cuda.synchronize = lambda *a, **k: None
    torch.cuda.Event = _Event
    torch.cuda.current_stream = lambda *a, **k: _Stream()
    torch.Tensor.pin_memory = lambda self, *a, **k: self
    bench.WORKLOADS["dry"] = ("tiny", 3, 64, 96, 8)               # T = 3 over 2 ranks: ragged frame shards
    sys.argv = ["bench.py", "--gpus", str(world), "--workload", "dry", "--steps", "1", "--warmup", "1", "--no-graph",
                "--precision", "fp32"]
    buf = io.StringIO()
    with oracle_ops(), contextlib.redirect_stdout(buf):
        bench.main()
    with open(os.path.join(out_dir, f"rank{rank}.out"), "w") as f:
        f.write(buf.getvalue())


def test_frame_sharded_run_control_flow_over_gloo(tmp_path):
    """the multi-GPU bench (driver: torchrun, one rank per GPU, NCCL) with two CPU processes over gloo: barrier + max-over-
    ranks timing, frame sharding + all-gather inside the step, only rank 0 prints, n_gpus / parallelism fields"""
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    out0, out1 = (open(tmp_path / f"rank{r}.out").read() for r in (0, 1))
    assert out1.strip() == ""
    lines = [l for l in out0.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for k in CONTRACT:
        assert k in line, k
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and line["value"] > 0
    assert line["config"]["parallelism"].startswith("frame-shard x2") and line["config"]["execution"] == "eager"
    assert line["cpu_baseline"] is None                    # measured at N = 1 only
    assert line["e2e"]["value"] > 0


@pytest.mark.parametrize("workload", ["dry_sot", "dry_grounding"])
def test_prompt_workload_control_flow(monkeypatch, workload):
    """BASELINE configs[2] / configs[3] (visual prompts with a grown memory pool, text prompts + lang->vision)"""
    monkeypatch.setattr(sys, "argv", ["bench.py", "--workload", workload, "--steps", "1", "--warmup", "1", "--precision", "fp32"])
    # the parser's choices were fixed at import: let it see the dry workloads
    line = _run(monkeypatch, ["--workload", workload, "--steps", "1", "--warmup", "1", "--precision", "fp32"])
    for k in CONTRACT:
        assert k in line, k
    assert line["value"] > 0 and line["config"]["execution"] == "eager" and workload in line["config"]["workload"]
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 64 * 96
    assert ("P=10" if workload == "dry_sot" else "P=32") in line["config"]["workload"]
