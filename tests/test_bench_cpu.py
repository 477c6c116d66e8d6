"""CPU: the host logic of bench.py that a GPU-less box can exercise -- where the roofline's peak and traffic
come from, the contig layout of the workload, the isolated child of the call-stage sample (it must hand back
an object whatever happens to it), the host re-packing rate, and the JSON keys of the contract named in the
source.  No kernel runs here."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from svdss_b200 import capi, synth  # noqa: E402


def test_contig_offsets_cover_the_reference():
    o = bench.contig_offsets(bench.REF_BP, bench.N_CONTIGS)
    assert o[0] == 0 and o[-1] == bench.REF_BP and len(o) == bench.N_CONTIGS + 1
    d = np.diff(o)
    assert (d > 0).all() and (np.diff(d) <= 1).all()          # GRCh38-like: contigs shrink from first to last
    assert 3.5 < d[0] / d[-1] < 4.5


def test_hbm_peak_prefers_the_measured_file(tmp_path, monkeypatch):
    monkeypatch.setattr(bench, "ROOT", str(tmp_path))
    v, src = bench.hbm_peak()
    assert v == bench.HBM_FALLBACK_GBS and src.startswith("fallback")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"hbm_gbs": 6551.0, "bf16_tflops": 1686.7}))
    v, src = bench.hbm_peak()
    assert v == 6551.0 and src.startswith("measured")
    (tmp_path / "MEASURED_PEAKS.json").write_text(json.dumps({"peaks": {"hbm_copy_tb_s": 6.4, "hbm_pct": 83}}))
    v, src = bench.hbm_peak()
    assert v == 6400.0 and "hbm_copy_tb_s" in src
    (tmp_path / "MEASURED_PEAKS.json").write_text("not json")
    assert bench.hbm_peak()[0] == bench.HBM_FALLBACK_GBS


def test_traffic_comes_from_the_committed_capture(monkeypatch):
    monkeypatch.delenv("SVB_SEARCH_CFG", raising=False)
    with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
        t = json.load(f)
    full = bench.ncu_traffic(1_000_000, 128, "k_sfs_search_mop")
    assert full == t["k_sfs_search_mop"]["dram_bytes"]
    assert bench.ncu_traffic(250_000, 128, "k_sfs_search_mop") == full / 4          # scaled per read
    assert bench.ncu_traffic(1_000_000, 64, "k_sfs_search_mop") is None             # another index layout: not that capture
    monkeypatch.setenv("SVB_SEARCH_CFG", "cpa")
    assert bench.ncu_traffic(1_000_000, 128, "k_sfs_search_mop") is None            # another kernel: not that capture
    monkeypatch.delenv("SVB_SEARCH_CFG")
    assert bench.ncu_traffic(1_000_000, 128, "no such kernel") is None


def test_call_stage_child_always_returns_an_object():
    r = bench.call_stage_child(0, limit_s=120)
    assert isinstance(r, dict)
    if capi.lib().svb_device_count() < 1:
        assert "no CUDA device" in r["error"] and "no CPU fallback" in r["error"]   # the product path fails loudly
    else:
        assert r["poa"]["clusters"] > 0 and r["ksw2"]["pairs"] > 0 and r["poa"]["cpu_oracle"]["identical_consensus"]
    r = bench.call_stage_child(0, limit_s=0.01)                                     # a hang costs the time limit, not the line
    assert "did not finish" in r["error"]


def test_host_pack2_rate_on_a_small_batch():
    contigs = synth.make_reference(50_000, seed=3, contigs=1)
    reads = synth.make_reads(contigs, 200, seed=4, mean_len=3000, sd_len=500, min_len=500, max_len=6000)
    reads[7] = reads[7].copy()
    reads[7][10] = 5
    seq4, s4o, lq = capi.pack_bam4(reads)
    r = bench.host_pack2_rate(capi, seq4, s4o, lq, n=150)
    with_n = sum(bool((x == 5).any()) for x in reads[:150])                         # reads that cannot travel as 2 bits per base
    assert with_n >= 1
    assert r["GB_s_of_4bit_input"] > 0 and r["sample_bytes"] == int(s4o[150] - s4o[0]) and r["reads_with_other_codes"] == with_n


def test_sv_scoring_and_host_threads(monkeypatch):
    truth = {(0, False, 300, 1000), (1, True, 80, 5000)}
    tab = np.array([[0, 0, 1001, 300], [1, 1, 5060, 80], [1, 1, 9000, 80]], np.int32)
    s = bench.score_calls(tab, truth)
    assert s == {"planted_with_two_carriers": 2, "recovered_exact_type_and_length": 2, "records": 3, "records_matching_no_planted_sv": 1}
    monkeypatch.setenv("OMP_NUM_THREADS", "1")          # what torchrun exports: the CPU arms must not shrink to one thread
    assert bench.host_threads() == len(os.sched_getaffinity(0))


def test_the_line_carries_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "workload", "e2e", "h2d_bytes_per_step", "d2h_bytes_per_step", "gpu_launches", "roofline",
                "bound", "achieved", "peak", "frac", "traffic", "cpu_baseline", "cores", "kind", "sample", "clocks", "impl"):
        assert '"%s"' % key in src, key
    assert "(search+call)" in src
