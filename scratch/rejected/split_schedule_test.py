"""CPU test of the split tournament schedule (rejected design, see scratch/rejected/README.md)."""
JB = 16

def rr_pair(n, r, p):
    if p == 0:
        return n - 1, r
    return (r + p) % (n - 1), (r - p + n - 1) % (n - 1)


def split_sweep(nblk):
    """The launches of one sweep of the split tournament (svd.cu launch_split_sweep), per stream: a list of stages,
    each a list of rounds, each a list of block pairs."""
    assert nblk % 4 == 0
    q, h = nblk // 4, nblk // 2

    def sched1(gA, gN, r):
        return [tuple(sorted((gA + a, gA + b))) for a, b in (rr_pair(gN, r, p) for p in range(gN // 2))]

    def sched2(gA, gB, gN, r):
        return [(gA + p, gB + (p + r) % gN) for p in range(gN)]

    streams = []
    for t in range(2):
        stage0 = [("diag", sched1(t * h, h, 0))] + [("cross", sched1(t * h, h, r)) for r in range(h - 1)]
        later = [[("cross", sched2(t * q, h + (q if (t ^ stage) else 0), q, r)) for r in range(q)] for stage in range(2)]
        streams.append([stage0] + later)
    return streams


def test_split_tournament_schedule():
    """Every block pair meets exactly once per sweep, every block is rotated in-block once, and the two streams
    never touch the same block inside a stage (they only synchronise at stage boundaries)."""
    for nblk in (8, 12, 64, 68):
        s0, s1 = split_sweep(nblk)
        cross, diag = [], []
        for stage0, stage1 in zip(s0, s1):
            blocks = [set(), set()]
            for t, stage in enumerate((stage0, stage1)):
                for kind, pairs in stage:
                    assert len(pairs) == nblk // 4
                    flat = [b for pr in pairs for b in pr]
                    assert len(set(flat)) == len(flat)          # pairs of a launch are disjoint
                    blocks[t].update(flat)
                    (diag if kind == "diag" else cross).extend(pairs)
            assert not (blocks[0] & blocks[1])
        assert sorted(cross) == [(a, b) for a in range(nblk) for b in range(a + 1, nblk)]
        assert sorted(b for pr in diag for b in pr) == list(range(nblk))
        assert sum(len(st) for st in s0) == nblk                # launches per stream and sweep
