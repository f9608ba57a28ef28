"""Pure-Python model of the CUDA kernel's algorithm (scrooge_b200/csrc/sg_align.cuh): left-aligned vectors,
G-row chunks with a forefront, V/H/E edge words instead of stored R rows, traceback over the edge words.
It exists so that the algorithmic identities the kernel relies on can be checked against the oracle on a
machine without a GPU (tests/test_kernel_model.py); it is not used by the product."""
from typing import List, Tuple

OPS = "=XID"
CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "a": 0, "c": 1, "g": 2, "t": 3}


def align(text: str, read: str, W: int = 64, O: int = 33, G: int = 8) -> Tuple[int, str, int]:
    MASK = (1 << W) - 1
    TBL = W - O
    TOP_SHIFT = W - 32
    t = [CODE[c] for c in text]
    q = [CODE[c] for c in read]
    t_pos = q_pos = 0
    ed = 0
    out: List[str] = []
    while q_pos < len(q):
        n = min(W, len(t) - t_pos)
        m = min(W, len(q) - q_pos)
        hm = (MASK << (W - m)) & MASK
        pm = [0, 0, 0, 0]
        for c in range(4):
            v = 0
            for J in range(m):
                if q[q_pos + J] != c:
                    v |= 1 << (W - 1 - J)
            pm[c] = v & hm
        FF = [0] * (W + 1)
        V = [0] * (TBL + 1)
        H = [0] * (TBL + 1)
        E = [0] * (TBL + 1)
        d0 = 0
        while True:
            first = d0 == 0
            C = [0] * G
            S = [0] * G
            XFp = MASK
            for i in range(n, -1, -1):
                F = MASK if first else FF[i]
                sF = (F << 1) & MASK
                v = h = e = 0
                if i == n:
                    for r in range(G):
                        s = W - m + d0 + r
                        C[r] = (MASK << s) & MASK if s < W else 0
                        S[r] = (C[r] << 1) & MASK
                        v |= C[r] & ~S[r]
                else:
                    p = pm[t[t_pos + i]]
                    e = p
                    aboveS = MASK if first else sF
                    aboveX = XFp
                    for r in range(G):
                        Xr = C[r] & S[r]
                        newC = ((S[r] | p) & aboveX) & aboveS
                        h |= newC & ~C[r]
                        C[r] = newC
                        aboveX = Xr
                        S[r] = (C[r] << 1) & MASK
                        v |= C[r] & ~S[r]
                        aboveS = S[r]
                XFp = MASK if first else (F & sF)
                FF[i] = C[G - 1]
                if i <= TBL:
                    top = lambda x: (x & MASK) >> TOP_SHIFT
                    V[i] = top(v) if first else V[i] | top(v)
                    H[i] = top(h) if first else H[i] | top(h)
                    if first:
                        E[i] = top(e)
            above = sum((C[r] >> (W - 1)) & 1 for r in range(G))
            if above == G:
                d0 += G
                continue
            break
        i = j = 0
        mask = 1 << 31
        cur_op, cur_cnt = None, 0
        while j < m and i < TBL and j < TBL:
            if V[i] & mask:
                op = 2
            elif H[i] & mask:
                op = 3
            elif E[i] & mask:
                op = 1
            else:
                op = 0
            if op != 2:
                i += 1
            if op != 3:
                j += 1
                mask >>= 1
            if op != 0:
                ed += 1
            if op != cur_op:
                if cur_cnt:
                    out.append(f"{cur_cnt}{OPS[cur_op]}")
                cur_op, cur_cnt = op, 1
            else:
                cur_cnt += 1
        if cur_cnt:
            out.append(f"{cur_cnt}{OPS[cur_op]}")
        t_pos += i
        q_pos += j
    return ed, "".join(out), t_pos


def align_delta(text: str, read: str, W: int = 64, O: int = 33) -> Tuple[int, str, int, int]:
    """Model of the delta-encoded DC (sg_align_delta.cuh): the same window matrix D(i,J) -- the one the R rows of
    src/genasm_cpu.cpp:210-288 encode as D(i,J) = min{d : bit J of R[d][i] is 0} -- computed column by column as
    vertical/horizontal +-1 deltas (Myers 1999 / Hyyro 2001 bit-vector recurrences run from column n down to 0 on
    left-aligned vectors), instead of row by row as K+1 threshold vectors.  V_i = vertical +1 deltas of column i,
    H_i = horizontal +1 deltas between columns i+1 and i: the very words the row-wise kernel ORs together.
    Returns (edit distance, cigar, consumed reference prefix, sum over windows of (d_w+1)(n+1))."""
    MASK = (1 << W) - 1
    TBL = W - O
    TOP_SHIFT = W - 32
    t = [CODE[c] for c in text]
    q = [CODE[c] for c in read]
    t_pos = q_pos = 0
    ed = 0
    entries = 0
    out: List[str] = []
    while q_pos < len(q):
        n = min(W, len(t) - t_pos)
        m = min(W, len(q) - q_pos)
        hm = (MASK << (W - m)) & MASK
        pm = [0, 0, 0, 0]
        for c in range(4):
            v = 0
            for J in range(m):
                if q[q_pos + J] != c:
                    v |= 1 << (W - 1 - J)
            pm[c] = v & hm
        V = [0] * (TBL + 1)
        H = [0] * (TBL + 1)
        E = [0] * (TBL + 1)
        top = lambda x: (x & MASK) >> TOP_SHIFT
        Pv, Mv = hm, 0          # boundary column D(n,J) = m-J: every vertical delta is +1
        if n <= TBL:
            V[n] = top(Pv)
        for i in range(n - 1, -1, -1):
            p = pm[t[t_pos + i]]
            Eq = ~p & MASK      # padding bits (below W-m) match everything: zero rows stay zero rows
            D0 = ((((Eq & Pv) + Pv) & MASK) ^ Pv) | Eq | Mv
            Ph = Mv | (~(D0 | Pv) & MASK)
            Mh = Pv & D0
            Phs = (Ph << 1) & MASK   # carry-in 0: D(i,m) = 0 for every i
            Mhs = (Mh << 1) & MASK
            Pv = Mhs | (~(D0 | Phs) & MASK)
            Mv = Phs & D0
            if i <= TBL:
                V[i], H[i], E[i] = top(Pv), top(Ph), top(p)
        d_w = bin(Pv).count("1") - bin(Mv).count("1")   # D(0,0) = sum of the vertical deltas of column 0
        entries += (d_w + 1) * (n + 1)
        i = j = 0
        mask = 1 << 31
        cur_op, cur_cnt = None, 0
        while j < m and i < TBL and j < TBL:
            if V[i] & mask:
                op = 2
            elif H[i] & mask:
                op = 3
            elif E[i] & mask:
                op = 1
            else:
                op = 0
            if op != 2:
                i += 1
            if op != 3:
                j += 1
                mask >>= 1
            if op != 0:
                ed += 1
            if op != cur_op:
                if cur_cnt:
                    out.append(f"{cur_cnt}{OPS[cur_op]}")
                cur_op, cur_cnt = op, 1
            else:
                cur_cnt += 1
        if cur_cnt:
            out.append(f"{cur_cnt}{OPS[cur_op]}")
        t_pos += i
        q_pos += j
    return ed, "".join(out), t_pos, entries


def align_delta_generic(text: str, read: str, W: int, O: int) -> Tuple[int, str, int, int]:
    """Model of the run-time (W, O) kernel (sg_align_generic.cuh): the delta recurrence of align_delta on vectors of
    WP = 32 * ceil(W / 32) bits (pattern position J at bit WP-1-J, so a window narrower than the vector is just a window
    with more padding rows), always W columns per window with the "matches nothing" mask standing in for the columns
    i >= n, the op planes A = V | H, B = ~V & (H | E) kept at full width for the W-O+1 traceback columns, and run-length
    encoding during the walk.  Returns (edit distance, cigar, consumed reference prefix, sum of (d_w+1)(n+1))."""
    NW = (W + 31) // 32
    WP = 32 * NW
    MASK = (1 << WP) - 1
    TBL = W - O
    t = [CODE[c] for c in text]
    q = [CODE[c] for c in read]
    t_pos = q_pos = 0
    ed = 0
    entries = 0
    out: List[str] = []
    while q_pos < len(q):
        n = min(W, len(t) - t_pos)
        m = min(W, len(q) - q_pos)
        hm = (MASK << (WP - m)) & MASK
        pm = [0, 0, 0, 0, hm]           # pm[4]: a character that matches nothing
        for c in range(4):
            v = 0
            for J in range(m):
                if q[q_pos + J] != c:
                    v |= 1 << (WP - 1 - J)
            pm[c] = v & hm
        A = [0] * (TBL + 1)
        B = [0] * (TBL + 1)
        Pv, Mv = hm, 0
        for i in range(W - 1, -1, -1):
            p = pm[t[t_pos + i]] if i < n else pm[4]
            Eq = ~p & MASK
            x = ((((Eq & Pv) + Pv) & MASK) ^ Pv) | Eq
            Ph = Mv | (~(x | Pv) & MASK)
            Mh = Pv & x
            Phs = (Ph << 1) & MASK
            Mhs = (Mh << 1) & MASK
            xv = Eq | Mv
            Pv = Mhs | (~(xv | Phs) & MASK)
            Mv = Phs & xv
            if i <= TBL:
                A[i] = Pv | Ph
                B[i] = ~Pv & (Ph | p) & MASK
        d_w = bin(Pv).count("1") - bin(Mv).count("1")
        entries += (d_w + 1) * (n + 1)
        i = j = 0
        jmax = min(m, TBL)
        # the walk writes the two bits of step k to bit k of two streams of four 32-bit words (at most 2 TBL <= 126 steps),
        # eight words when W - O > 63 (at most 254 steps)
        SW = 8 if TBL > 63 else 4
        hs, ls, k = [0] * SW, [0] * SW, 0
        while j < jmax and i < TBL:
            bit = 1 << (WP - 1 - j)
            hi, lo = bool(A[i] & bit), bool(B[i] & bit)
            if hi:
                hs[k >> 5] |= 1 << (k & 31)
            if lo:
                ls[k >> 5] |= 1 << (k & 31)
            k += 1
            if not (hi and not lo):
                i += 1
            if not (hi and lo):
                j += 1
        runs, edits = rle_streams(hs, ls, j)
        assert sum(c for c, _ in runs) == k
        ed += edits
        out.extend(f"{c}{OPS[o]}" for c, o in runs)
        t_pos += i
        q_pos += j
    return ed, "".join(out), t_pos, entries


def rle_streams(hs: List[int], ls: List[int], j: int):
    """The kernels' run-length encoding on the op streams, word by word with their bit tricks: run boundaries
    hs ^ hs >> 1 | ls ^ ls >> 1 (the streams are zero beyond the last step), the last step forced, steps = j + #D,
    edits = popc(hs | ls); one (count, op) per set boundary bit, the op read at the boundary's position."""
    SW = len(hs)
    M32 = 0xFFFFFFFF
    steps = j + sum(bin(h & l).count("1") for h, l in zip(hs, ls))
    edits = sum(bin(h | l).count("1") for h, l in zip(hs, ls))
    e = []
    for w in range(SW):
        hn = ((hs[w] >> 1) | ((hs[w + 1] if w + 1 < SW else 0) << 31)) & M32
        ln = ((ls[w] >> 1) | ((ls[w + 1] if w + 1 < SW else 0) << 31)) & M32
        e.append((hs[w] ^ hn) | (ls[w] ^ ln))
    for w in range(SW):
        last = steps - 1 - 32 * w
        if 0 <= last < 32:
            e[w] |= 1 << last
        if last < 31:
            e[w] &= 0 if last < 0 else (2 << last) - 1
    runs = []
    st = -1
    for w in range(SW):
        ew = e[w]
        while ew:
            p = (ew & -ew).bit_length() - 1
            op = ((hs[w] >> p) & 1) * 2 + ((ls[w] >> p) & 1)
            runs.append((p - st, op))
            st = p
            ew &= ew - 1
        st -= 32
    return runs, edits


def walk_checked(A: List[int], B: List[int], m: int, TBL: int):
    """The traceback walk with its end tests (reference src/genasm_cpu.cpp:307-310): op planes A/B per column, pattern
    position j at bit 31-j, op = 2A+B (0 '=', 1 'X', 2 'I', 3 'D').  Returns (ops, i, j)."""
    i = j = 0
    ops: List[int] = []
    jmax = min(m, TBL)
    while j < jmax and i < TBL and len(ops) < 2 * TBL:
        op = 2 * ((A[i] >> (31 - j)) & 1) + ((B[i] >> (31 - j)) & 1)
        ops.append(op)
        if op != 2:
            i += 1
        if op != 3:
            j += 1
    return ops, i, j


def walk_unchecked_then_cut(A: List[int], B: List[int], m: int, TBL: int):
    """What genasm_delta_kernel does (SG_DELTA_SHORTFAST): TB_LIMIT steps without any end test -- also for a window with
    m < TB_LIMIT pattern characters, whose steps beyond j == m read whatever the planes hold below the pattern -- then,
    for such a window, the op streams are cut at the m-th step that consumed a pattern character (every op but 'D'),
    found by halving with popcounts; a walk that is not over after TB_LIMIT steps goes on with the end tests."""
    i = j = 0
    h = l = 0
    for k in range(TBL):
        a, b = (A[i] >> (31 - j)) & 1, (B[i] >> (31 - j)) & 1
        h |= a << k
        l |= b << k
        if not (a and not b):
            i += 1
        if not (a and b):
            j += 1
    steps = TBL
    if m < TBL:
        walked = (1 << TBL) - 1
        nd = ~(h & l) & walked
        if bin(nd).count("1") >= m:
            pos, r = 0, m
            for half in (16, 8, 4, 2, 1):
                c = bin((nd >> pos) & ((1 << half) - 1)).count("1")
                if c < r:
                    r -= c
                    pos += half
            keep = (2 << pos) - 1
            h &= keep
            l &= keep
            i = bin(~(h & ~l) & keep).count("1")
            j = m
            steps = pos + 1
            return [2 * ((h >> k) & 1) + ((l >> k) & 1) for k in range(steps)], i, j
    ops = [2 * ((h >> k) & 1) + ((l >> k) & 1) for k in range(steps)]
    jmax = min(m, TBL)
    while j < jmax and i < TBL and len(ops) < 2 * TBL:
        op = 2 * ((A[i] >> (31 - j)) & 1) + ((B[i] >> (31 - j)) & 1)
        ops.append(op)
        if op != 2:
            i += 1
        if op != 3:
            j += 1
    return ops, i, j
