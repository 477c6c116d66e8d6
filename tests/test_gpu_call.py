"""GPU: `SVDSS call --bam --sfs` end to end (Clusterer on the host, POA + ksw2 on the GPU, CIGAR walk,
clean_dups / filter_sv_chains, VCF) against the Python restatement over the oracle, byte for byte, on a
diploid sample with planted INS/DEL; the planted SVs must come back with exact type and length."""
import os
import subprocess

import pytest

import call_model
import cluster_model
from sv_world import make_world
from svdss_b200 import build

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def exe():
    build.build_lib()
    return build.build_host()


def vcf_body(text):
    return [l for l in text.splitlines() if not l.startswith("#")]


def check_planted(world, body):
    found = 0
    for sv in world["catalogue"]:
        chrom = world["names"][sv["contig"]]
        for l in body:
            f = l.split("\t")
            if f[0] == chrom and abs(int(f[1]) - (sv["pos"] + 1)) <= 20 and ("SVTYPE=%s;" % sv["type"]) in f[7] \
                    and ("SVLEN=%d;" % (sv["len"] if sv["type"] == "INS" else -sv["len"])) in f[7]:
                found += 1
                break
    return found


@pytest.mark.parametrize("tag_hp,noht,threads", [(True, False, 4), (True, True, 3), (False, False, 2)])
def test_call_from_bam_and_sfs(exe, tmp_path, tag_hp, noht, threads):
    w = make_world(str(tmp_path), tag_hp=tag_hp, seed=71 + threads)
    args = [exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--threads", str(threads),
            "--poa", str(tmp_path / "poa.sam")]
    if noht:
        args.append("--noht")
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    body = vcf_body(r.stdout)
    # the same with the Clusterer's BAM scan on the device (svb_bamstream_*, alignment mode): whole file in one window, and
    # windows that cut records
    for window in (None, "5000"):
        env = dict(os.environ, SVB_BGZF_GPU_MIN_BYTES="0")
        if window:
            env["SVB_BGZF_WINDOW"] = window
        g = subprocess.run(args + ["--gpu-inflate"], capture_output=True, text=True, env=env)
        assert g.returncode == 0, g.stderr
        assert "BAM records decoded on GPU" in g.stderr
        assert g.stdout == r.stdout, window
    clusters = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=threads)
    want = call_model.call_vcf_lines(w["ref_seqs"], clusters, useht=not noht, threads=threads)
    assert body == want
    assert check_planted(w, body) >= 8
    if tag_hp and not noht:
        assert any("COV0=-1" in l for l in body)        # a haplotype sub-cluster made a call
    assert all("RVEC=1:" in l or "RVEC=0:" in l for l in body)
    sam = (tmp_path / "poa.sam").read_text().splitlines()
    assert sam[0] == "@HD\tVN:1.4" and sum(1 for l in sam if not l.startswith("@")) >= len(body) // 2
