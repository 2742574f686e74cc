"""The range-addressable input generator of config 4 (tools/lac_synth.c: lac_synth_range).

Any rank must be able to produce its block range of the ONE 10 h file: a range generated on its own has
to equal the same frames cut out of a whole-file generation, for every depth / channel layout, for starts
that are and are not reset points, and the no-reset form has to be the original Appendix C stream."""
import numpy as np
import pytest

import helpers as H

SS = 1 << 19  # super-section: the four 2^17-frame sections once


@pytest.mark.parametrize("depth,channels", [(24, 2), (16, 2), (24, 1)])
def test_range_equals_slice_of_whole(depth, channels):
    total = 3 * SS + 70001
    wl, wr, wp = H.synth_range(4, 0, total, depth, channels, want_packed=True)
    fb = channels * (depth // 8)
    for f0, n in [(0, 1000), (SS, 50000), (SS - 7, 4000), (2 * SS + 12345, SS + 999), (16384 * 33, 16384 * 5 + 3),
                  (total - 5, 5), (3 * SS, 70001)]:
        l, r, p = H.synth_range(4, f0, n, depth, channels, want_packed=True)
        assert np.array_equal(l, wl[f0:f0 + n])
        if channels == 2:
            assert np.array_equal(r, wr[f0:f0 + n])
        assert np.array_equal(p, wp[f0 * fb:(f0 + n) * fb])


def test_no_reset_form_is_the_appendix_c_stream():
    total = SS + 5000
    l0, r0, p0 = H.synth(2, total, 24, want_packed=True)
    l1, r1, p1 = H.synth_range(2, 0, total, 24, 2, reset_log2=0, want_packed=True)
    assert np.array_equal(l0, l1) and np.array_equal(r0, r1) and np.array_equal(p0, p1)
    # a no-reset range away from 0 runs the recurrences from frame 0 and is still the same stream
    l2, r2 = H.synth_range(2, SS - 100, 3000, 24, 2, reset_log2=0)
    assert np.array_equal(l2, l0[SS - 100:SS + 2900]) and np.array_equal(r2, r0[SS - 100:SS + 2900])


def test_reset_stream_matches_appendix_c_below_one_super_section():
    l0, r0 = H.synth(4, SS, 24)
    l1, r1 = H.synth_range(4, 0, SS, 24)
    assert np.array_equal(l0, l1) and np.array_equal(r0, r1)
    # ... and differs after the first reset (the filter state restarts), so the two specs are distinct
    l0, _ = H.synth(4, SS + 4096, 24)
    l1, _ = H.synth_range(4, 0, SS + 4096, 24)
    assert not np.array_equal(l0[SS:], l1[SS:])


def test_sections_still_hit_every_path():
    """second super-section of the reset stream: AR noise, triangle, sparse silence, level-stepped noise"""
    l, r = H.synth_range(4, SS, SS, 24)
    sec = [l[i << 17:(i + 1) << 17] for i in range(4)]
    assert np.abs(sec[0]).max() > 1 << 18 and np.abs(np.diff(sec[0].astype(np.int64))).mean() > 1000
    assert np.abs(sec[1]).max() <= (12100 << 8) + 255
    assert (sec[2] == 0).mean() > 0.9
    assert np.abs(sec[3]).max() > 1 << 22
