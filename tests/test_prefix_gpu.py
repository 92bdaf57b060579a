"""K9: the parallel running sum must be bit-identical to the reference's
sequential loop `c += w[i]` (particle_filter.cpp:225-236, particle.cpp:66-82).

numpy's cumsum over float64 is that loop (np.add.accumulate is strictly
sequential), so it is the oracle here; the library's one-warp sequential kernel
is the second witness.  Comparison is on the raw 64-bit patterns."""
import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def filt():
    occ, res, origin = syn.make_map(400)
    gm = q.create_likelihood_map(device=0)
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    return q.create_particle_filter(gm, seed=1)


def seq_cumsum(w, carry=0.0):
    return np.cumsum(np.concatenate([[carry], w]))[1:]


def check(filt, w, carry=0.0, oracle=True):
    w = np.ascontiguousarray(w, np.float64)
    filt.set_prefix_mode(0)          # parallel, predicted binade crossings summarised
    got, co = filt.debug_prefix(w, carry)
    filt.set_prefix_mode(2)          # parallel, every binade-crossing chunk by literal additions
    got2, co2 = filt.debug_prefix(w, carry)
    filt.set_prefix_mode(1)          # the literal one-warp loop
    seq, co_seq = filt.debug_prefix(w, carry)
    filt.set_prefix_mode(0)
    assert np.array_equal(got2.view(np.uint64), seq.view(np.uint64)), "mode 2 != sequential"
    assert np.float64(co2).view(np.uint64) == np.float64(co_seq).view(np.uint64)
    assert np.array_equal(got.view(np.uint64), seq.view(np.uint64)), \
        f"parallel != one-warp sequential at {np.flatnonzero(got.view(np.uint64) != seq.view(np.uint64))[:5]}"
    assert np.float64(co).view(np.uint64) == np.float64(co_seq).view(np.uint64)
    if oracle:
        with np.errstate(all="ignore"):
            want = seq_cumsum(w, carry)
        assert np.array_equal(got.view(np.uint64), want.view(np.uint64)), \
            f"first mismatch at {np.flatnonzero(got.view(np.uint64) != want.view(np.uint64))[:5]}"
        if len(w):
            assert co == want[-1] or (np.isnan(co) and np.isnan(want[-1]))
    return got


@pytest.mark.parametrize("n", [1, 7, 31, 32, 255, 256, 257, 2047, 2048, 2049, 4096, 100_000,
                               1_000_003])
def test_equal_weights(filt, n):
    check(filt, np.full(n, 1.0 / n))


@pytest.mark.parametrize("n,seed", [(1000, 1), (50_000, 2), (300_000, 3), (2_000_000, 4)])
def test_random_normalised_weights(filt, n, seed):
    rng = np.random.default_rng(seed)
    w = rng.lognormal(0.0, 3.0, n)
    w[rng.random(n) < 0.2] = 0.0           # penalised particles
    w /= w.sum()
    check(filt, w)
    # the parallel sum is NOT what a reassociated sum gives: make sure the test can tell
    pair = w.reshape(-1, 2).sum(axis=1).cumsum() if n % 2 == 0 else None
    if pair is not None and n >= 300_000:
        assert not np.array_equal(pair, seq_cumsum(w)[1::2])


def test_ties_round_half_even(filt):
    # carry in [0.5, 1): ulp = 2^-53; weights that are odd multiples of 2^-54 are exact ties
    rng = np.random.default_rng(5)
    n = 200_000
    odd = (2 * rng.integers(0, 1 << 20, n) + 1).astype(np.float64)
    w = odd * 2.0 ** -54
    w[::3] = rng.random(len(w[::3])) * 2.0 ** -34   # non-ties in between
    w[5::17] = 2.0 ** -54                           # the smallest tie: q = 0
    got = check(filt, w, carry=0.5)
    assert got[-1] < 1.0
    # and in the binade of a sum that starts from nothing
    check(filt, w)


def test_binade_crossings_everywhere(filt):
    # the sum doubles every ~50 elements for 600 binades
    i = np.arange(30_000)
    w = 2.0 ** (i / 50.0 - 620.0)
    check(filt, w)
    # one dominant weight in the middle, dust before and after
    rng = np.random.default_rng(6)
    w = rng.random(100_000) * 1e-12
    w[43_210] = 1.0
    check(filt, w)
    # descending magnitudes: later weights are absorbed completely
    w = 2.0 ** -(np.arange(5000) / 10.0)
    check(filt, w)


def test_carry_in_chains_shards(filt):
    rng = np.random.default_rng(7)
    w = rng.random(700_001)
    w /= w.sum()
    whole = check(filt, w)
    cut = [0, 1, 255, 100_000, 350_000, 700_001]
    carry = 0.0
    for a, b in zip(cut[:-1], cut[1:]):
        part = check(filt, w[a:b], carry)
        assert np.array_equal(part, whole[a:b])
        carry = part[-1]


def test_degenerate_inputs_fall_back_to_true_additions(filt):
    rng = np.random.default_rng(8)
    w = rng.random(20_000)
    w[100] = -0.25            # negative
    w[5000] = 5e-324          # subnormal
    w[12_000] = 0.0
    w[12_001] = -0.0
    check(filt, w)
    w[15_000] = np.inf
    check(filt, w)
    w[15_000] = np.nan
    check(filt, w, oracle=False)   # NaN payloads: compare the two kernels only
    check(filt, np.zeros(3000))
    check(filt, np.zeros(0))
    z = np.zeros(10_000)
    z[7777] = 1.0
    check(filt, z)


ULP = 2.0 ** -53      # of the binade [0.5, 1)


@pytest.mark.parametrize("seed", [11, 12])
def test_crossing_at_every_element_of_a_chunk(filt, seed):
    # carry 0.5 + dust, then one weight that takes the sum across 1.0 at element j of chunk 3,
    # for every j; ties on both sides of the crossing (odd multiples of ulp/2 before it, of ulp
    # after it), zeros sprinkled in
    rng = np.random.default_rng(seed)
    for j in list(range(0, 256, 5)) + [1, 254, 255]:
        n = 2048 + int(rng.integers(0, 300))
        w = (2 * rng.integers(0, 1 << 12, n) + 1).astype(np.float64) * (ULP / 2)
        w[rng.random(n) < 0.1] = 0.0
        at = 3 * 256 + j
        w[at] = 0.5 - w[:at].sum() + float(rng.integers(0, 3)) * ULP
        w[at + 1:] = (2 * rng.integers(0, 1 << 12, n - at - 1) + 1).astype(np.float64) * ULP
        w[at + 1::7] = 0.0
        got = check(filt, w, carry=0.5)
        assert got[at - 1] < 1.0 <= got[at] or got[at] < 1.0    # the crossing is at `at` or just after


def test_mispredicted_crossings_fall_back(filt):
    # (a) dust that the sequential sum absorbs but a tree sum accumulates: the approximate carry
    # runs ahead, the crossing is predicted too early
    dust = np.full(200_000, 2.0 ** -55)
    near = np.array([0.5 - 1000 * ULP])
    steps = np.full(40, 100 * ULP)
    w = np.concatenate([dust, near, steps, np.full(3000, 2.0 ** -30)])
    got = check(filt, w, carry=0.5)
    assert got[len(dust)] < 1.0 and got[len(dust) + 11] >= 1.0
    # (b) every addition rounds up (0.75 ulp -> 1 ulp): the approximate carry lags, the crossing
    # comes earlier than predicted
    up = np.full(100_000, 0.75 * ULP)
    near = np.array([0.5 - 90_000 * ULP])
    w = np.concatenate([up, near, np.full(4000, 1000 * ULP)])
    got = check(filt, w, carry=0.5)
    assert got[len(up) - 1] < 1.0 <= got[len(up)]
    # (c) two crossings inside one chunk, and a crossing in the very last (partial) chunk
    w = np.concatenate([np.full(300, 2.0 ** -20), [0.5, 1.0], np.full(100, 2.0 ** -20)])
    check(filt, w, carry=0.5)
    w = np.concatenate([np.full(2048 + 17, 2.0 ** -20), [0.5 - 1000 * 2.0 ** -20], np.full(5, 2.0 ** -20)])
    check(filt, w, carry=0.5)


def test_crossings_with_the_largest_exponents(filt):
    # the binade above the carry's must exist: sums near the top of the double range
    w = np.full(5000, 2.0 ** 1010)
    check(filt, w, carry=2.0 ** 1022)
    w = np.full(5000, 2.0 ** 1000)
    check(filt, w, carry=2.0 ** 1021)
