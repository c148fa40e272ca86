"""Host logic of the band partition (pfb_imaging_b200/split.py): LPT assignment and the plane-offload schedule."""
from pfb_imaging_b200 import split as bs

# per-band ms / planes of the C2 job measured on a B200 (profiles/r1_bench_c2_n1_v11_final.json)
C2_MS = [8.452, 9.445, 10.038, 11.134, 11.867, 12.634, 13.793, 14.507]
C2_P = [8, 9, 9, 10, 10, 10, 11, 11]


def _tplane():
    return [0.47 * ms / p for ms, p in zip(C2_MS, C2_P)]  # ~47 % of an apply is plane transforms


def test_lpt_assign_balances_and_is_deterministic():
    for world in (1, 2, 4, 8):
        own = bs.lpt_assign(C2_MS, world)
        assert own == bs.lpt_assign(list(C2_MS), world)
        loads = [sum(ms for ms, o in zip(C2_MS, own) if o == r) for r in range(world)]
        assert set(own) == set(range(world))
        assert max(loads) <= sum(C2_MS) / world * (1.03 if world < 8 else 1.3)


def test_offload_schedule_beats_the_one_band_per_gpu_bound():
    own = bs.lpt_assign(C2_MS, 8)
    off, loads = bs.plan_offloads(C2_MS, _tplane(), C2_P, own, 8)
    assert off, "the heavy bands must shed planes"
    # modelled step time: the 1 -> 8 speed-up of the job rises from sum/max = 6.33 to above 7
    assert sum(C2_MS) / max(loads) > 7.0 > sum(C2_MS) / max(C2_MS)
    for b, o in off.items():
        assert o["owner"] == own[b] and o["helper"] != o["owner"] and 1 <= o["nq"] < C2_P[b]
    # heaviest band is among the offloaded ones, helpers are lighter than owners
    assert 7 in off and C2_MS[own.index(off[7]["helper"])] < C2_MS[7]
    # nothing to do on one GPU or when the partition is already balanced
    assert bs.plan_offloads(C2_MS, _tplane(), C2_P, [0] * 8, 1)[0] == {}
    own4 = bs.lpt_assign(C2_MS, 4)
    off4, loads4 = bs.plan_offloads(C2_MS, _tplane(), C2_P, own4, 4)
    base4 = max(sum(ms for ms, o in zip(C2_MS, own4) if o == r) for r in range(4))
    assert max(loads4) <= base4 + 1e-9


def test_offload_schedule_never_makes_things_worse():
    import random

    rnd = random.Random(3)
    for _ in range(50):
        nb, world = rnd.randint(1, 12), rnd.randint(1, 8)
        ms = [rnd.uniform(2, 20) for _ in range(nb)]
        P = [rnd.randint(1, 16) for _ in range(nb)]
        tp = [0.5 * m / p for m, p in zip(ms, P)]
        own = bs.lpt_assign(ms, world)
        base = [0.0] * world
        for m, o in zip(ms, own):
            base[o] += m
        off, loads = bs.plan_offloads(ms, tp, P, own, world)
        assert max(loads) <= max(base) + 1e-9
        for b, o in off.items():
            assert 1 <= o["nq"] <= max(1, P[b] - 1) and o["helper"] != o["owner"]
