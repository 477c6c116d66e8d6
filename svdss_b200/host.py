"""Host-side mirror of the reference's interface for the search path (same names, argument meaning
and error behaviour), over the C ABI.  The compute is libsvdss_b200; nothing here falls back to a
CPU implementation.

  SFS                    sfs.hpp:31-79
  parse_sfsfile          sfs.cpp:5-30
  Assembler.assemble     assembler.cpp:34-56   (host version, used only for already-parsed .sfs data;
                                                the search path assembles on the device)
  PingPong.process_batch ping_pong.cpp:176-209 (whole batch at once instead of one thread slot)
  PingPong.output_batch  ping_pong.cpp:213-236 (.sfs text, byte-compatible incl. the trailing TAB)
  PingPong.search        ping_pong.cpp:239-397 (FASTX mode only: BAM input is the C++ shell's job, svdss_b200/host/io.hpp + svdss_main.cpp)
"""
import gzip
import sys
from collections import OrderedDict

import numpy as np

from . import capi

# ping_pong.hpp:46-52 (seq_nt6_table; rb3_char2nt6 has the same mapping)
NT6 = np.full(256, 5, np.uint8)
NT6[0] = 0
for _c, _v in (("A", 1), ("C", 2), ("G", 3), ("T", 4)):
    NT6[ord(_c)] = _v
    NT6[ord(_c.lower())] = _v


class SFS:
    """sfs.hpp:31-79 (fields kept name for name)"""
    __slots__ = ("chrom", "qname", "rs", "re", "qs", "qe", "l", "htag")

    def __init__(self, qname, qs, l, htag, chrom="", rs=0, re=0):
        self.chrom = chrom
        self.qname = qname
        self.rs, self.re = rs, re
        self.qs = qs
        self.qe = qs + l
        self.l = l
        self.htag = htag

    def __lt__(self, o):  # sfs.hpp:67-74
        if self.chrom == "" or o.chrom == "":
            return self.qs < o.qs
        return self.rs < o.rs if self.chrom == o.chrom else self.chrom < o.chrom

    def __eq__(self, o):  # sfs.hpp:76-78
        return self.chrom == o.chrom and self.rs == o.rs and self.re == o.re

    def __repr__(self):
        return "SFS(%s,%d,%d,%d)" % (self.qname, self.qs, self.l, self.htag)


def parse_sfsfile(path):
    """sfs.cpp:5-30: up to four whitespace tokens per line, '*' repeats the previous read name."""
    out = OrderedDict()
    name = None
    with open(path) as f:
        for line in f:
            info = line.split()[:4]
            if not info:
                continue
            if info[0] != "*":
                name = info[0]
                out[name] = []
            out[name].append(SFS(name, int(info[1]), int(info[2]), int(info[3])))
    return out


class Assembler:
    @staticmethod
    def assemble(sfs):
        """assembler.cpp:34-56"""
        sfs = sorted(sfs)
        out = []
        i = 0
        while i < len(sfs):
            j = i + 1
            while j < len(sfs) and sfs[j - 1].qs + sfs[j - 1].l > sfs[j].qs:
                j += 1
            l = sfs[j - 1].qs + sfs[j - 1].l - sfs[i].qs
            out.append(SFS(sfs[i].qname, sfs[i].qs, l, sfs[i].htag))
            i = j
        return out


def read_fastx(path):
    """kseq-style FASTA/FASTQ reader (fastq.hpp / kseq.h): yields (name, sequence bytes)."""
    op = gzip.open if str(path).endswith(".gz") else open
    with op(path, "rb") as f:
        name, seq, state = None, [], 0
        for raw in f:
            line = raw.rstrip(b"\r\n")
            if state == 2:          # FASTQ quality lines: consume as many bytes as the sequence had
                qleft -= len(line)
                if qleft <= 0:
                    state = 0
                continue
            if (line[:1] == b"@" and state == 0) or (line[:1] == b">" and state in (0, 1)):
                if name is not None:
                    yield name, b"".join(seq)
                name = line[1:].split()[0].decode() if len(line) > 1 else ""
                seq = []
                state = 1
            elif line[:1] == b"+" and state == 1 and name is not None:
                qleft = sum(len(s) for s in seq)
                state = 2 if qleft > 0 else 0
            elif state == 1:
                seq.append(line)
        if name is not None:
            yield name, b"".join(seq)


class PingPong:
    """Mirror of class PingPong (ping_pong.hpp:55-91) for the search path."""

    def __init__(self, index, assemble=True, putative=True, overlap=-1, bsize=200_000, out=None):
        self.index = index            # capi.Index  (rb3_fmi_t in the reference)
        self.assemble = assemble      # !--noassemble
        self.putative = putative      # !--noputative (only meaningful in BAM mode: XF tag filter)
        self.overlap = overlap        # config.hpp:82
        self.bsize = bsize            # --bsize; the GPU wants batches far larger than the reference's 10000
        self.out = out or sys.stdout
        self.reads_processed = 0
        self.total_sfs = 0

    def process_batch(self, names, seqs_nt6, htags=None):
        """ping_pong.cpp:176-209 for a whole batch. Returns OrderedDict qname -> [SFS] in the
        reference's per-read order (ascending qs when assembled, emit order otherwise). Reads that
        yield nothing still get an (empty) entry, like `solutions[qname]` does."""
        cat, offs = _concat(seqs_nt6)
        res = self.index.sfs_batch(cat, offs, overlap=self.overlap, assemble=self.assemble)
        out = OrderedDict()
        for r, name in enumerate(names):
            ht = htags[r] if htags is not None else 0
            a, b = int(res.offs[r]), int(res.offs[r + 1])
            out[name] = [SFS(name, int(res.qs[i]), int(res.len[i]), ht) for i in range(a, b)]
        self.reads_processed += len(names)
        return out

    def output_batch(self, batch):
        """ping_pong.cpp:213-236: `<qname or *>\\t<qs>\\t<l>\\t<htag>\\t\\n`"""
        w = self.out.write
        for name, sfss in batch.items():
            first = True
            for s in sfss:
                w("%s\t%d\t%d\t%d\t\n" % (name if first else "*", s.qs, s.l, s.htag))
                first = False
            self.total_sfs += len(sfss)

    def search_fastx(self, path):
        """ping_pong.cpp:239-397 in FASTX mode (htag 0, no XF filter)."""
        names, seqs = [], []
        for name, seq in read_fastx(path):
            names.append(name)
            seqs.append(NT6[np.frombuffer(seq, np.uint8)])
            if len(names) == self.bsize:
                self.output_batch(self.process_batch(names, seqs))
                names, seqs = [], []
        if names:
            self.output_batch(self.process_batch(names, seqs))
        return 0


def _concat(seqs):
    offs = np.zeros(len(seqs) + 1, np.int64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.concatenate(seqs).astype(np.uint8) if seqs and offs[-1] else np.zeros(0, np.uint8)
    return np.ascontiguousarray(cat), offs
