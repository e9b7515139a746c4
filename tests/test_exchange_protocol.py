"""Barrier protocol of the blind-rotation kernel's exchange buffers, checked as a model (CPU).

compute-sanitizer's racecheck is not available on the GPU pool, so the hazard analysis written next to
`fwd_transform`, `inv_transform_pair` and the end-of-step barrier in zig-tfhe_b200/csrc/blind_rotate.cu is restated
here as data and checked mechanically.  A ciphertext is owned by two warps that run the same straight-line program
and meet at named group barriers (`bar.sync`, which also orders their shared-memory accesses).  Two accesses by
different warps are ordered iff a barrier lies between them, i.e. iff they fall into different barrier epochs; so
the protocol is race-free iff, in every epoch, no warp writes a region the other warp reads or writes.

Regions: (buffer, rows of warp r, columns of warp c) for the two X2 buffers F and G -- slot = 73 row + 9 j + k with row and k
in 0..7 (`x2_slot` / `x1a_slot` in negacyclic_fft.cuh); rows 0-3 and columns k = 0-3 carry warp 0's index `hi`, 4-7 warp 1's --
and (accumulator, positions owned by warp w) for the two halves of the TRLWE accumulator.  "own" = this warp's rows / columns /
positions only, "all" = both warps' (the cross-warp side of an exchange, or the rotated reads of `load_rot_diffs`).

The sequences below transcribe the production configuration (six ciphertexts per CTA, X1 laid over the X2 buffers,
double-buffered X2, paired inverse transforms, no end-of-step barrier) for L digits per polynomial, plus the
negative controls that show the checker sees the hazards the barriers are there for."""
import pytest

BAR = ("bar",)


def W(buf, rows, cols="all"):
    return ("w", buf, rows, cols)


def R(buf, rows, cols="all"):
    return ("r", buf, rows, cols)


def forward_transform(flip):
    """fwd_transform<DBX2, ALIAS>: X1 in this warp's rows of the OTHER buffer, X2 through buffer `flip`"""
    other = 1 - flip
    return [W(("x", other), "own"), R(("x", other), "own"),      # X1: write, __syncwarp, read (same warp)
            W(("x", flip), "all", "own"),                          # X2 write x2_slot(lo, q, hi): every row, this warp's columns
            BAR,
            R(("x", flip), "own")], other                          # X2 read x2_slot(hi, lo, q): own rows, every column; flip toggles


def inverse_pair(flip, bar3=True):
    """inv_transform_pair: a through F = flip, b through G"""
    F, G = ("x", flip), ("x", 1 - flip)
    seq = [W(F, "own"), BAR,                                       # a: X2 write (own rows); bar 1
           R(F, "all", "own"), W(G, "own"), BAR,                   # a: X2 read x2_slot(lo, q, hi) (all rows, own columns); b: X2 write; bar 2
           R(G, "all", "own"),                                     # b: X2 read
           W(F, "own"), R(F, "own"),                               # a: X1 over F
           W(("acc", "a"), "own")]                                 # fin(a): reductions on own positions
    if bar3:
        seq.append(BAR)                                            # bar 3
    seq += [W(G, "own"), R(G, "own"),                              # b: X1 over G
            W(("acc", "b"), "own")]                                # fin(b)
    return seq, flip


def inverse_sequential(flip):
    """inv_transform<DBX2, ALIAS> twice (the pre-pairing kernel), each: X2 write own, bar, read all, bar, X1 over the same buffer"""
    seq = []
    for half in "ab":
        buf = ("x", flip)
        seq += [W(buf, "own"), BAR, R(buf, "all", "own"), BAR, W(buf, "own"), R(buf, "own"), W(("acc", half), "own")]
        flip = 1 - flip
    return seq, flip


def step(flip, L, paired=True, end_barrier=False, bar3=True):
    seq = []
    for half in "ab":
        seq.append(R(("acc", half), "all"))                        # load_rot_diffs: rotated reads touch any position
        for _ in range(L):
            s, flip = forward_transform(flip)
            seq += s
    s, flip = inverse_pair(flip, bar3) if paired else inverse_sequential(flip)
    seq += s
    if end_barrier:
        seq.append(BAR)
    return seq, flip


def races(program):
    """program: the per-warp access sequence (both warps run it).  Returns the list of (epoch, region) conflicts."""
    epochs, cur = [], []
    for ev in program:
        if ev == BAR:
            epochs.append(cur)
            cur = []
        else:
            cur.append(ev)
    epochs.append(cur)
    found = []
    for e, evs in enumerate(epochs):
        def regions(w, kinds):
            out = set()
            for kind, buf, rows, cols in evs:
                if kind in kinds:
                    rr = (w,) if rows == "own" else (0, 1)
                    cc = (w,) if cols == "own" else (0, 1)
                    out |= {(buf, r, c) for r in rr for c in cc}
            return out
        for w in (0, 1):
            clash = regions(w, "w") & regions(1 - w, "rw")
            found += [(e, r) for r in sorted(clash, key=str)]
    return found


def run(L, steps=4, **kw):
    prog, flip = [BAR], 0        # the barrier after the accumulator is initialised
    for _ in range(steps):
        s, flip = step(flip, L, **kw)
        prog += s
    return races(prog)


@pytest.mark.parametrize("L", [1, 2, 3, 4])
def test_production_protocol_is_race_free(L):
    """paired inverse, no end-of-step barrier: nine group barriers per step at L = 3"""
    assert run(L) == []
    prog, _ = step(0, L)
    assert prog.count(BAR) == 2 * L + 3


@pytest.mark.parametrize("L", [1, 3])
def test_sequential_inverse_is_race_free_with_and_without_the_end_of_step_barrier(L):
    """the pre-pairing kernel (two single inverse transforms, 2 L + 5 barriers with the end-of-step one): each add-back is
    followed by at least one group barrier before the matching rotated reads, so that barrier was redundant there too"""
    assert run(L, paired=False, end_barrier=True) == []
    assert run(L, paired=False, end_barrier=False) == []


def test_forward_transform_needs_its_barrier():
    prog, flip = [BAR], 0
    for _ in range(4):
        s, flip = forward_transform(flip)
        prog += [e for e in s if e != BAR]
    assert races(prog)


def nonalias_step(flip, L, end_barrier):
    """KCT <= 4 kernels: X1 has a buffer of its own (own rows only), X2 double-buffered; single inverse transforms"""
    seq = []
    x1 = ("x1", 0)
    for half in "ab":
        seq.append(R(("acc", half), "all"))
        for _ in range(L):
            seq += [W(x1, "own"), R(x1, "own"), W(("x", flip), "all", "own"), BAR, R(("x", flip), "own")]
            flip = 1 - flip
    for half in "ab":
        seq += [W(("x", flip), "own"), BAR, R(("x", flip), "all", "own"), W(x1, "own"), R(x1, "own"), W(("acc", half), "own")]
        flip = 1 - flip
    if end_barrier:
        seq.append(BAR)
    return seq, flip


def exact_step(flip, L, end_barrier):
    """exact kernel, six per CTA: every transform (both directions) runs A -> X1 -> B -> X2 -> C like fwd_transform"""
    seq = []
    for half in "ab":
        seq.append(R(("acc", half), "all"))
        for _ in range(L):
            s, flip = forward_transform(flip)
            seq += s
    for half in "ab":
        s, flip = forward_transform(flip)
        seq += s + [W(("acc", half), "own")]
    if end_barrier:
        seq.append(BAR)
    return seq, flip


@pytest.mark.parametrize("stepfn", [nonalias_step, exact_step])
@pytest.mark.parametrize("L", [1, 2, 3])
def test_other_kernels_need_no_end_of_step_barrier(stepfn, L):
    """each half's add-back is followed by a group barrier (the other half's transform, or the next step's first digit transform)
    before the rotated reads that depend on it, and X2 is double-buffered: the barrier at the end of a step is redundant"""
    for end_barrier in (True, False):
        prog, flip = [BAR], 0
        for _ in range(4):
            s, flip = stepfn(flip, L, end_barrier)
            prog += s
        assert races(prog) == []


@pytest.mark.parametrize("L", [1, 3])
def test_every_remaining_barrier_is_needed(L):
    """minimality: dropping any one of the 2 L + 3 group barriers of a step (in every step) produces a conflict"""
    for k in range(2 * L + 3):
        prog, flip = [BAR], 0
        for _ in range(4):
            s, flip = step(flip, L)
            seen, kept = 0, []
            for ev in s:
                if ev == BAR:
                    if seen != k:
                        kept.append(ev)
                    seen += 1
                else:
                    kept.append(ev)
            prog += kept
        assert races(prog), f"barrier {k} of a step would be redundant"


def test_third_barrier_of_the_pair_is_needed():
    """bar 3 orders everybody's X2 reads of G before b's X1 reuses G -- and a's add-back before the next step's rotated reads"""
    bad = run(3, bar3=False)
    assert any(r[0] == ("x", 0) or r[0] == ("x", 1) for _, r in bad)
    assert any(r[0] == ("acc", "a") for _, r in bad)


def test_forward_x2_write_without_double_buffering_races():
    """control: with a single X2 buffer a transform's cross-warp X2 writes meet the previous transform's reads (why X2 is double-buffered)"""
    prog = [BAR]
    for _ in range(3):
        prog += [W(("x", 0), "all", "own"), BAR, R(("x", 0), "own")]
    assert races(prog)
