"""GPU: `SVDSS call --clipped` end to end (caller.cpp:37-55): the VCF body is what `call` prints without
the flag, followed by the Clipper's records for the clips of this run and the regions of the SVs just
called -- the same records the `_clipper` hook gives for those inputs (checked against the Python
transcription in tests/test_clipper_cpu.py)."""
import os
import subprocess

import pytest

from test_clipper_cpu import clip_world, exe  # noqa: F401  (fixtures)

pytestmark = pytest.mark.gpu


def test_call_clipped_appends_the_clipper_records(exe, clip_world):
    w = clip_world
    base = [exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--threads", "4"]
    plain = subprocess.run(base, capture_output=True, text=True)
    assert plain.returncode == 0, plain.stderr
    clips = os.path.join(w["d"], "gpu_clips.tsv")
    clipped = subprocess.run(base + ["--clipped", "--clips", clips], capture_output=True, text=True)
    assert clipped.returncode == 0, clipped.stderr
    assert clipped.stdout.startswith(plain.stdout)
    # clips come out of the same records when the BAM scan runs on the device
    dev = subprocess.run(base + ["--clipped", "--clips", clips + ".dev", "--gpu-inflate"], capture_output=True, text=True,
                         env=dict(os.environ, SVB_BGZF_GPU_MIN_BYTES="0", SVB_BGZF_WINDOW="7000"))
    assert dev.returncode == 0, dev.stderr
    assert "BAM records decoded on GPU" in dev.stderr and dev.stdout == clipped.stdout
    assert open(clips + ".dev").read() == open(clips).read()
    tail = clipped.stdout[len(plain.stdout):].splitlines()
    body = [l.split("\t") for l in plain.stdout.splitlines() if not l.startswith("#")]
    regions = os.path.join(w["d"], "gpu_regions.tsv")
    with open(regions, "w") as f:
        for t in body:
            end = int([kv for kv in t[7].split(";") if kv.startswith("END=")][0][4:])
            f.write("%d %d\n" % (int(t[1]) - 1000, end + 1000))
    hook = subprocess.run([exe, "_clipper", "--reference", w["fa"], "--clips-in", clips, "--regions-in", regions, "--threads", "4"],
                          capture_output=True, text=True)
    assert hook.returncode == 0, hook.stderr
    assert tail == hook.stdout.splitlines()
    assert "Predicted %d SVs from clipped alignments" % len(tail) in clipped.stderr
