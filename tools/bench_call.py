"""Measurement of the `call` kernels at the SURVEY 8(d) config-4 (POA) and config-5 (ksw2) shapes.
  python tools/bench_call.py [--clusters N] [--pairs N] [--max-len L]
Prints one JSON line per kernel: throughput, GCUPS, and the CPU oracle timed on a bounded sample."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import oracle
from svdss_b200 import build, capi


from svdss_b200.synth import gen_clusters, gen_pairs  # noqa: E402  (SURVEY 8d config 4 / 5 shapes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clusters", type=int, default=2000)
    ap.add_argument("--pairs", type=int, default=20000)
    ap.add_argument("--max-len", type=int, default=10000)
    ap.add_argument("--cpu-seconds", type=float, default=8.0)
    a = ap.parse_args()
    build.build_lib()
    rng = np.random.default_rng(6)
    if a.clusters:
        cl = gen_clusters(rng, a.clusters)
        capi.poa_batch(cl[:64])
        res = capi.poa_batch(cl)
        t0 = time.perf_counter(); k = 0; ok = True
        while time.perf_counter() - t0 < a.cpu_seconds and k < len(cl):
            ok &= bool(np.array_equal(oracle.poa_consensus(cl[k], band=True), res.consensus(k))); k += 1
        cpu = k / (time.perf_counter() - t0)
        print(json.dumps({"kernel": "k_poa", "clusters": len(cl), "reads": sum(len(c) for c in cl),
                          "kernel_ms": round(res.kernel_ms, 1), "device_ms": round(res.device_ms, 1),
                          "clusters_per_s": round(len(cl) / res.device_ms * 1e3, 1), "cells": res.cells,
                          "GCUPS": round(res.cells / res.kernel_ms / 1e6, 3), "reruns": res.reruns,
                          "cpu_oracle_clusters_per_s_1core": round(cpu, 2), "parity_sample": ok, "sample": k}), flush=True)
    if a.pairs:
        pr = gen_pairs(rng, a.pairs, hi=a.max_len)
        qc, qo = oracle.concat([p[0] for p in pr]); tc, to = oracle.concat([p[1] for p in pr])
        capi.ksw_extd2_batch(qc[:qo[64]], qo[:65], tc[:to[64]], to[:65])
        res = capi.ksw_extd2_batch(qc, qo, tc, to)
        t0 = time.perf_counter(); k = 0; ok = True; cells = 0
        while time.perf_counter() - t0 < a.cpu_seconds and k < len(pr):
            sc, cg = oracle.ksw_extd2(*pr[k])
            ok &= (sc == int(res.score[k]) and cg == res.cigar_of(k)); cells += len(pr[k][0]) * len(pr[k][1]); k += 1
        dt = time.perf_counter() - t0
        print(json.dumps({"kernel": "k_ksw_extd2", "pairs": len(pr), "cells": res.cells, "waves": res.waves,
                          "kernel_ms": round(res.kernel_ms, 1), "device_ms": round(res.device_ms, 1),
                          "pairs_per_s": round(len(pr) / res.device_ms * 1e3, 1),
                          "GCUPS": round(res.cells / res.kernel_ms / 1e6, 2),
                          "cpu_oracle_GCUPS_1core": round(cells / dt / 1e9, 3), "parity_sample": ok, "sample": k}), flush=True)


if __name__ == "__main__":
    main()
