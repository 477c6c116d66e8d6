"""Pure-Python literal transcription of the reference's search hot path over a *naive bidirectional
FMD index* -- oracle statement #3 (small inputs only).

Follows, line by line:
  ping_pong.cpp:4-49   PingPong::ping_pong_search
  SURVEY A.1 / 8(a2)   rb3_fmd_set_intv, rb3_fmd_extend (ropebwt3 @0ea3919, restated; the source is
                       not vendored in the reference tree)
and adds the index-free definition (`occurs()` is a plain substring test over contigs and their
reverse complements) so that the three statements can be compared in tests/test_oracle.py.
"""
import numpy as np


def comp6(c):
    return 5 - c if 1 <= c <= 4 else c


def build_text(contigs):
    T = []
    for s in contigs:
        s = list(int(x) for x in s)
        T += s + [0]
        T += [comp6(c) for c in reversed(s)] + [0]
    return T


class NaiveFMD:
    """BWT by sorting suffixes (sentinels ordered by position), full Occ table, bi-intervals."""

    def __init__(self, contigs):
        T = build_text(contigs)
        n = len(T)
        self.n = n
        # distinct sentinels: replace each 0 by a unique negative key ordered by position
        key = [(c if c else -(n - i)) for i, c in enumerate(T)]
        # -(n-i): earlier position -> more negative -> sorts first
        sa = sorted(range(n), key=lambda i: key[i:])
        self.sa = sa
        bwt = [T[i - 1] if i else T[n - 1] for i in sa]
        self.bwt = bwt
        self.acc = [0] * 7
        for c in T:
            self.acc[c + 1] += 1
        for c in range(6):
            self.acc[c + 1] += self.acc[c]
        occ = np.zeros((n + 1, 6), np.int64)
        for i, c in enumerate(bwt):
            occ[i + 1] = occ[i]
            occ[i + 1][c] += 1
        self.occ = occ

    def set_intv(self, c):
        # rb3_fmd_set_intv: x[0]=acc[c], size=acc[c+1]-acc[c], x[1]=acc[comp(c)]
        return [self.acc[c], self.acc[comp6(c)], self.acc[c + 1] - self.acc[c]]

    def extend(self, ik, is_back):
        # rb3_fmd_extend(f, ik, ok[6], is_back)
        x0, x1, size = ik
        x = [x0, x1]
        nb = 0 if is_back else 1  # index !is_back
        ib = 1 if is_back else 0
        tk = self.occ[x[nb]]
        tl = self.occ[x[nb] + size]
        ok = [[0, 0, 0] for _ in range(6)]
        for c in range(6):
            ok[c][nb] = self.acc[c] + int(tk[c])
            ok[c][2] = int(tl[c] - tk[c])
        ok[0][ib] = x[ib]
        ok[4][ib] = ok[0][ib] + ok[0][2]
        ok[3][ib] = ok[4][ib] + ok[4][2]
        ok[2][ib] = ok[3][ib] + ok[3][2]
        ok[1][ib] = ok[2][ib] + ok[2][2]
        ok[5][ib] = ok[1][ib] + ok[1][2]
        return ok


def ping_pong_search(index, P, overlap=-1):
    """ping_pong.cpp:4-49; P is a list of nt6 codes with P[l] == 0 appended like ping_pong.cpp:94."""
    l = len(P) - 1
    solutions = []
    begin = l - 1
    while begin >= 0:
        ik = index.set_intv(P[begin])
        while ik[2] != 0 and begin > 0:
            begin -= 1
            ok = index.extend(ik, 1)
            ik = ok[P[begin]]
        if begin == 0 and ik[2] != 0:
            break
        end = begin
        ik = index.set_intv(P[end])
        while ik[2] != 0:
            end += 1
            ok = index.extend(ik, 0)
            c = P[end]
            ik = ok[5 - c if 1 <= c <= 4 else c]
        solutions.append((begin, end - begin + 1))
        if begin == 0:
            break
        if overlap == 0:
            begin -= 1
        else:
            begin = end + overlap
    return solutions


def occurs_factory(contigs):
    strands = []
    for s in contigs:
        s = bytes(int(x) for x in s)
        strands.append(s)
        strands.append(bytes(comp6(c) for c in reversed(s)))

    def occurs(w):
        w = bytes(w)
        return any(w in s for s in strands)

    return occurs


def sfs_definition(contigs, P):
    """Index-free statement: brute-force substring tests only."""
    occ = occurs_factory(contigs)
    P = [int(x) for x in P]
    l = len(P)
    out = []
    s = l - 1
    while s >= 0:
        b = s
        while b >= 0 and occ(P[b:s + 1]):
            b -= 1
        if b < 0:
            break
        e = b
        while occ(P[b:e + 1]):
            e += 1
        out.append((b, e - b + 1))
        if b == 0:
            break
        s = e - 1
    return out


def assemble(pairs):
    """assembler.cpp:34-56"""
    sfs = sorted(pairs)
    out = []
    i = 0
    while i < len(sfs):
        j = i + 1
        broke = False
        while j < len(sfs):
            if sfs[j - 1][0] + sfs[j - 1][1] <= sfs[j][0]:
                out.append((sfs[i][0], sfs[j - 1][0] + sfs[j - 1][1] - sfs[i][0]))
                i = j
                broke = True
                break
            j += 1
        if not broke:
            out.append((sfs[i][0], sfs[j - 1][0] + sfs[j - 1][1] - sfs[i][0]))
            i = j
    return out
