"""CPU checks of bench.py's own plumbing (no GPU, no timing): the pieces of the JSON line that are built from constants and recorded runs,
the one-line-on-stdout guard, and the operation counts of the roofline -- a formatting slip there costs a whole GPU run."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_roofline_notes_and_recorded_runs_are_well_formed():
    note = bench.oz_traffic_note(7, 50048)
    assert "7 planes" in note and "18.8 GB" in note and "16 % of the HBM peak" in note
    assert bench.OZ_LAUUM_TRAFFIC_BYTES is None or bench.OZ_LAUUM_TRAFFIC_BYTES > 1e11
    fit = bench.recorded_fits()
    assert fit is None or (fit["recorded"] is True and all(r["gradient_calls"] > 0 and r["wall_s"] > 0 for r in fit["runs"]))
    json.dumps({"note": note, "fit": fit})
    # int8 op counts: the B^-1 launch is part of the evaluation's total, which is just below n^3 per slice pair
    n_pad = 50048
    assert 0 < bench.int8_macs_lauum(n_pad) < bench.int8_macs(n_pad) < float(n_pad) ** 3


def test_stdout_carries_exactly_the_json_line():
    """Libraries print to file descriptor 1 (NCCL's version banner): after guard_stdout() only emit() reaches the real stdout."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.guard_stdout(); print('banner from python'); "
            "os.write(1, b'banner from C\\n'); bench.emit({'metric': 'm', 'value': 1.5})" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
    assert "banner from python" in out.stderr and "banner from C" in out.stderr


class _FakeModel:
    """Stands in for gp_ss_ak_b200.GpssModel so that bench.main() can assemble its JSON line without a GPU (no arithmetic is checked here)."""
    slices = 7

    def __init__(self, X, y, device=0):
        self.n = X.shape[0]
        self.launches = 0

    def padded_n(self):
        return (self.n + 127) // 128 * 128

    def set_theta(self, th):
        pass

    def set_data(self, X, y):
        pass

    def nlml_grad(self):
        self.launches += 100
        import numpy as np
        return 1.0, np.ones(10)

    def last_call_ms(self):
        return 5.0

    def launch_count(self):
        return self.launches

    def set_profiling(self, on):
        pass

    def phase_ms(self):
        import numpy as np
        return np.array([0.1, 2.0, 0.3, 1.5, 1.2, 0.1, 0.2, 0.3, 0.0] + [0.0] * 7)

    def ozaki_slices(self):
        return _FakeModel.slices

    def ozaki_digit_bits(self):
        return 8

    def predict_shard(self, m_total, sums, shard):
        import numpy as np
        return np.zeros(len(shard)), np.ones(len(shard))

    def close(self):
        pass


import pytest  # noqa: E402


@pytest.mark.parametrize("slices", [7, 0])
def test_bench_main_assembles_its_line_on_both_pipes(monkeypatch, slices):
    """bench.main() end to end with the device replaced by a stub: every string of the roofline / predict / phases blocks is formatted, for the
    int8 pipe and for GPSS_OZAKI=0, and the line carries the keys the contract names."""
    import torch
    import gp_ss_ak_b200 as G
    _FakeModel.slices = slices
    lines = []
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(G, "GpssModel", _FakeModel)
    monkeypatch.setattr(G, "measure_fp64_peak", lambda d: 37.0, raising=False)
    monkeypatch.setattr(bench, "int8_tensor_peak", lambda d: (3800.0, 4100.0, "stub"))
    monkeypatch.setattr(bench, "guard_stdout", lambda: None)
    monkeypatch.setattr(bench, "emit", lines.append)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "1", "--warmup", "1", "--n", "2000", "--pred-m", "256", "--no-cpu-baseline"])
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)
    bench.main()
    assert len(lines) == 1
    line = json.loads(json.dumps(lines[0]))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                "config", "e2e", "gpu_launches", "clocks", "roofline", "phases_ms", "predict"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["scaling"] == "strong" and line["vs_baseline"] is None
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert ("int8" in line["dtype"]) == (slices > 0)
