"""The code path bench.py times, replayed and checked: `Engine(streams=8, ring=8)`, per step one `h2d` of the ring entry and, per
stream group, `encode_group` (staggered I / P phases: only one group is in an I frame at a time) + `d2h_group`, packed levels,
copies overlapping compute, nothing synchronised between steps.  Every slot's decisions and packed levels of every step, and the
final reconstructions, must equal the oracle's (bit-exact): a missing event wait between the copy-in, compute and copy-out
streams, a ring entry overwritten too early or a result set reused too early shows up here.  Contract: one result per picture,
in order (av_encode.c:970)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


class OracleSlot:
    """the oracle's state of one closed-GOP chain"""

    def __init__(self, oracle, w, h, prm):
        self.o, self.w, self.h, self.prm = oracle, w, h, prm
        self.prev = None; self.pmv = None

    def step(self, frame, is_i):
        o = self.o
        cur = o.OFrame(self.w, self.h).load(*frame); rec = o.OFrame(self.w, self.h)
        info, coef = o.encode_frame(self.prm, 0 if is_i else 1, cur, None if is_i else self.prev, rec, None if is_i else self.pmv)
        self.prev = rec
        self.pmv = np.zeros(info.size, o.MV); self.pmv["x"] = info["mvx"]; self.pmv["y"] = info["mvy"]
        return info, coef, rec


def replay(oracle, b2, w, h, slots, streams, ring, gop, steps, R=16, qp=28, deblock=0, check_slots=None, distinct=True, me_prune=0):
    eng = b2.Engine(w, h, slots=slots, fmt="yuv420p", ring=ring, merange=R, qp=qp, subpel=1, intra_in_p=1, streams=streams,
                    deblock=deblock, pack_levels=1, me_prune=me_prune)
    groups = eng.groups(); NG = len(groups)
    phase = [g * gop // NG for g in range(NG)]                       # bench.py: staggered GOP phases
    group_of = {s: g for g, (s0, n) in enumerate(groups) for s in range(s0, s0 + n)}
    check_slots = list(range(slots)) if check_slots is None else check_slots
    prm = oracle.Params(qp, R, 1, 1, deblock)
    chains = {s: OracleSlot(oracle, w, h, prm) for s in check_slots}
    n_y = w * h

    def is_i(step, g):
        return step == 0 or (step + phase[g]) % gop == 0

    def picture(step, s):
        return oracle.synth_frame(w, h, step if distinct else step % ring, s)

    def fill(step):
        for s in range(slots):
            y, u, v = picture(step, s)
            buf = eng.host_input(s, step % ring)
            buf[:n_y] = y.ravel(); buf[n_y:n_y + u.size] = u.ravel(); buf[n_y + u.size:] = v.ravel()

    def issue(step):                                                 # == bench.py issue(step, with_copies=True)
        r = step % ring
        eng.h2d(ring=r)
        sets = []
        for g in range(NG):
            eng.encode_group(g, b2.FRAME_I if is_i(step, g) else b2.FRAME_P, ring=r)
            eng.d2h_group(g)
            sets.append(eng.group_result_set(g))
        return sets

    def verify(step, sets):
        for g in range(NG):
            eng.group_wait(g, sets[g])                               # that copy only; later work keeps running on the GPU
        for s in check_slots:
            g = group_of[s]
            info_g, packed_g = eng.results_set(sets[g], s)
            info_o, coef_o, _ = chains[s].step(picture(step, s), is_i(step, g))
            want = b2.shipped_info(info_o)                            # the copy-out ships 24-byte decision records
            for f in info_o.dtype.names:
                assert np.array_equal(info_g[f], want[f]), f"step {step} slot {s} (group {g}, {'I' if is_i(step, g) else 'P'}): info.{f}"
            assert np.array_equal(packed_g, b2.pack_levels(info_o, coef_o)), f"step {step} slot {s}: packed levels"

    if not distinct:
        for r in range(ring):
            fill(r)
    pending = None
    for step in range(steps):
        if distinct:
            fill(step)                                               # fresh pictures every step: ring entry step % ring is reused
        sets = issue(step)
        if pending is not None:
            verify(*pending)                                         # step - 1 is checked while step runs
        pending = (step, sets)
    verify(*pending)
    eng.sync()
    for s in check_slots:
        ry, ru, rv = eng.recon(s)
        rec = chains[s].prev
        assert np.array_equal(ry, rec.y) and np.array_equal(ru, rec.u) and np.array_equal(rv, rec.v), f"final recon of slot {s}"
    launches = eng.launch_count()
    eng.close()
    return launches


@pytest.mark.parametrize("slots,deblock", [(16, 0), (24, 1)])
def test_bench_issue_path_every_slot_every_step(oracle, b2, slots, deblock):
    """8 stream groups, 8-deep ring, GOP 8 with staggered phases, 44 steps: every group crosses five GOP boundaries; every slot
    of every step is compared"""
    launches = replay(oracle, b2, 96, 64, slots=slots, streams=8, ring=8, gop=8, steps=44, deblock=deblock)
    assert launches > 44 * 8 * 8


def test_bench_issue_path_64_slots(oracle, b2):
    """bench.py's slot count (64 GOPs in 8 groups of 8), GOP 32, 40 steps: groups with phase >= 24 start a second GOP"""
    replay(oracle, b2, 64, 48, slots=64, streams=8, ring=8, gop=32, steps=40, R=32, qp=26)


def test_bench_issue_path_1080p_spot_check(oracle, b2):
    """C3 size (1920x1080, +-32, QP 26, 64 slots, ring pictures repeating like bench.py's): first slot of two groups over I, P, P"""
    replay(oracle, b2, 1920, 1080, slots=64, streams=8, ring=8, gop=32, steps=3, R=32, qp=26, check_slots=[0, 40], distinct=False)


def test_bench_issue_path_pruned_search(oracle, b2):
    """the same path with me_prune=1 (bench.py's `pruned` leg): K1a + pruned K1 of eight stream groups overlap, the block-sum planes
    are per slot and refilled every P step; every slot of every step still equals the oracle's exhaustive search.  Plus the C3-size
    spot check."""
    replay(oracle, b2, 96, 64, slots=16, streams=8, ring=8, gop=8, steps=28, R=32, deblock=1, me_prune=1)
    replay(oracle, b2, 1920, 1080, slots=64, streams=8, ring=8, gop=32, steps=3, R=32, qp=26, check_slots=[0, 40], distinct=False, me_prune=1)
