"""Row N4: the pre-filters hqdn3d and yadif.  libavfilter is not in the image, so these are pinned only by
the C restatement oracle/b2o_filters.c (GPU bit-exact, -m gpu) plus the properties the published algorithms guarantee
(CPU): a static progressive picture passes yadif unchanged, hqdn3d lowers temporal noise without moving the mean."""
import numpy as np
import pytest
from test_oracle_decode import smooth_seq, coarse_seq


def _interlaced(w, h, n, seed=0):
    """fields taken at different instants of a fast pan woven into frames (top field = earlier instant)"""
    prog = smooth_seq(w, 2 * h, 2 * n, seed=seed)            # double height so that fields keep detail
    frames = []
    for t in range(n):
        a, b = prog[2 * t], prog[2 * t + 1]
        def weave(pa, pb):
            f = pa.copy(); f[1::2] = pb[1::2]; return f
        y = weave(a[0][:h], b[0][:h]); u = weave(a[1][:(h + 1) // 2], b[1][:(h + 1) // 2]); v = weave(a[2][:(h + 1) // 2], b[2][:(h + 1) // 2])
        frames.append((np.ascontiguousarray(y), np.ascontiguousarray(u[:, :(w + 1) // 2]), np.ascontiguousarray(v[:, :(w + 1) // 2])))
    return frames


def test_oracle_yadif_on_static_progressive_content(oracle):
    """the kept field passes untouched; the rebuilt lines stay at the temporal prediction (= the original line) unless the
    vertical neighbourhood allows a small excursion (yadif's b/f check) -- and exactly there on a vertically flat picture"""
    w, h = 96, 64
    f = smooth_seq(w, h, 1, seed=3)[0]
    out = oracle.yadif_sequence([f, f, f], w, h)
    for o in out:
        assert np.array_equal(o[0][0::2], f[0][0::2])
        assert np.abs(o[0][1::2].astype(int) - f[0][1::2]).mean() < 2.0
    cols = np.tile(f[0][:1], (h, 1)); cu = np.tile(f[1][:1], (h // 2, 1))
    g = (cols, cu, cu)
    for o in oracle.yadif_sequence([g, g, g], w, h):
        assert np.array_equal(o[0], cols)


def test_oracle_yadif_removes_combing(oracle):
    w, h = 128, 96
    frames = _interlaced(w, h, 4, seed=5)
    out = oracle.yadif_sequence(frames, w, h, tff=1)
    def comb(y):                                              # energy between adjacent lines relative to lines two apart
        y = y.astype(int)
        return np.abs(y[1:-1] * 2 - y[:-2] - y[2:]).mean()
    assert comb(out[2][0]) < 0.6 * comb(frames[2][0])
    assert np.array_equal(out[2][0][0::2], frames[2][0][0::2])           # the kept (top) field is untouched


def test_oracle_hqdn3d_lowers_noise_keeps_mean(oracle):
    w, h, n = 96, 64, 8
    base = coarse_seq(w, h, 1, seed=7, scale=16, noise=0.0)[0]
    rng = np.random.default_rng(1)
    noisy = [tuple(np.clip(p.astype(int) + rng.integers(-6, 7, p.shape), 0, 255).astype(np.uint8) for p in base) for _ in range(n)]
    f = oracle.Hqdn3d(w, h)
    out = [f(fr) for fr in noisy]
    err_in = np.mean([(a[0].astype(int) - base[0]) ** 2 for a in noisy[3:]])
    err_out = np.mean([(a[0].astype(int) - base[0]) ** 2 for a in out[3:]])
    assert err_out < 0.6 * err_in
    assert abs(np.mean([a[0].mean() for a in out[3:]]) - base[0].mean()) < 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("fmt,w,h,spec,args", [("yuv420p", 176, 144, "hqdn3d", {}), ("yuv420p", 150, 98, "hqdn3d=2:1:3:1", dict(ls=2, cs=1, lt=3, ct=1)),
                                               ("yuv422p", 96, 70, "hqdn3d=6", dict(ls=6)), ("yuv411p", 128, 48, "hqdn3d=4:3:0:0", dict(ls=4, cs=3, lt=0, ct=0))])
def test_gpu_hqdn3d_matches_oracle(oracle, b2, fmt, w, h, spec, args):
    rng = np.random.default_rng(2)
    dims = oracle._plane_dims(fmt, w, h)
    seq = smooth_seq(w, h, 5, seed=11)
    frames = [tuple(np.clip(np.resize(fr[0 if p == 0 else 1], (ph, pw)).astype(int) + rng.integers(-5, 6, (ph, pw)), 0, 255).astype(np.uint8)
                    for p, (pw, ph) in enumerate(dims)) for fr in seq]
    g = b2.FilterGraph(w, h, spec, fmt=fmt)
    got = g.run(frames)
    g.close()
    ref = oracle.Hqdn3d(w, h, fmt, **args)
    assert len(got) == len(frames)
    for t, (fr, (planes, pts)) in enumerate(zip(frames, got)):
        exp = ref(fr)
        assert pts == 100 + t
        for p in range(3):
            assert np.array_equal(planes[p], exp[p]), f"hqdn3d frame {t} plane {p}"


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,tff,spec", [(176, 144, 1, "yadif"), (150, 98, 0, "yadif=0:1"), (64, 34, 1, "yadif=0:0"), (720, 576, 1, "yadif")])
def test_gpu_yadif_matches_oracle(oracle, b2, w, h, tff, spec):
    frames = _interlaced(w, h, 5, seed=w)
    g = b2.FilterGraph(w, h, spec)
    got = []
    for t, f in enumerate(frames):
        g.add(f, pts=100 + t, tff=tff)
        assert g.poll() == (1 if t >= 1 else 0)                # one frame of delay: yadif needs the next picture
        while g.poll() > 0:
            got.append(g.get())
    g.flush()
    assert g.poll() == 1
    got.append(g.get())
    assert g.get() is None
    g.close()
    exp = oracle.yadif_sequence(frames, w, h, tff=tff)
    assert [p for _, p in got] == [100 + t for t in range(5)]
    for t in range(5):
        for p in range(3):
            assert np.array_equal(got[t][0][p], exp[t][p]), f"yadif frame {t} plane {p}"


@pytest.mark.gpu
def test_gpu_filter_chain_and_errors(oracle, b2):
    """the author's DV chain "hqdn3d,yadif" (av_encode.c:35) equals the composition of the two oracles; pass-through graph;
    unknown filters are refused like avfilter_graph_parse fails (av_encode.c:499-502)"""
    w, h = 160, 120
    frames = _interlaced(w, h, 5, seed=21)
    g = b2.FilterGraph(w, h, "hqdn3d,yadif")
    got = g.run(frames)
    g.close()
    dn = oracle.Hqdn3d(w, h)
    exp = oracle.yadif_sequence([dn(f) for f in frames], w, h, tff=1)
    assert len(got) == 5
    for t in range(5):
        for p in range(3):
            assert np.array_equal(got[t][0][p], exp[t][p]), f"chain frame {t} plane {p}"
    g = b2.FilterGraph(w, h, "")
    out = g.run(frames[:2]); g.close()
    assert all(np.array_equal(a, b) for a, b in zip(out[1][0], frames[1]))
    with pytest.raises(RuntimeError):
        b2.FilterGraph(w, h, "unsharp")
    with pytest.raises(RuntimeError):
        b2.FilterGraph(w, h, "yadif=1")
