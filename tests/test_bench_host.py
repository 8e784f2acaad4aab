"""CPU checks of bench.py's host-side logic (no GPU, no CUDA library): the MEASURED_PEAKS.json reader."""
import importlib.util
import json
import os
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")


def _bench(monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py"])
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_hbm_peak_reader_accepts_the_drivers_shapes(tmp_path, monkeypatch):
    b = _bench(monkeypatch)
    monkeypatch.setattr(b, "ROOT", str(tmp_path))
    assert b.hbm_peak() == (b.FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)")
    cases = [({"hbm_gbs": 6541.5, "bf16_tflops": 1700.0}, 6541.5),
             ({"hbm": {"burst_gbs": 7000.0, "sustained_gbs": 6541.5}, "dense_bf16_tfs": 1600.0}, 6541.5),   # sustained wins
             ({"copy_bandwidth_tb_s": 6.5}, 6500.0),                                                          # TB/s
             ({"bf16_tflops": 1700.0}, b.FALLBACK_HBM_GBS)]                                                   # nothing recognised
    for doc, want in cases:
        (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps(doc))
        got, src = b.hbm_peak()
        assert got == want, (doc, got, src)
