"""CPU check of the arithmetic identity dc_chain_kernel relies on (icspcodec_b200/csrc/icsp_kernels.cuh): the reference's
double sequence for a DC level, L = (int)trunc|floor((raw - P) + 0.5) / Q (ENC:2780 luma, ENC:4642 chroma, after
DPCM_DC_block), equals an integer expression of m = floor(raw + 0.5) whenever raw + 0.5 is not within 2^-30 of an integer.
numpy float64 is IEEE binary64 with round-to-nearest-even, i.e. the arithmetic of the reference's doubles."""
import numpy as np

EPS = 2.0 ** -30


def ref_level(raw, P, Q, chroma):
    x = (raw - P.astype(np.float64)) + 0.5                 # two rounded double operations, like the reference
    r = np.floor(x) if chroma else np.trunc(x)
    r = r.astype(np.int64)
    return np.sign(r) * (np.abs(r) // Q)                   # C division: toward zero


def int_level(raw, P, Q, chroma):
    xs = raw + 0.5
    m = np.floor(xs)
    g = xs - m                                             # exact
    amb = ~((g >= EPS) & (g <= 1.0 - EPS)) | ~(np.abs(xs) < 8192.0)
    d = m.astype(np.int64) - P
    if not chroma:
        d = d + (d < 0)                                    # truncation toward zero of the non-integer (m - P) + g
    return np.sign(d) * (np.abs(d) // Q), amb


def _cases(rng, n):
    raw = rng.uniform(-2041.0, 2041.0, n)
    # the values a DCT of integers actually produces (multiples of 1/8 times 1 + 3e-8), exact integers, exact halves, and
    # values a few ulps / 1e-12 / 1e-9 / 1e-7 away from k + 0.5
    k = rng.integers(-2040, 2041, n).astype(np.float64)
    eighth = rng.integers(-16320, 16321, n) / 8.0 * (1.0 - 3.4e-8)
    near = k + 0.5 + rng.choice([0.0, 1e-13, -1e-13, 1e-12, -1e-12, 2e-9, -2e-9, 1e-7, -1e-7, 2.0 ** -31, -(2.0 ** -31)], n)
    ulps = np.nextafter(k + 0.5, rng.choice([-np.inf, np.inf], n))
    return np.concatenate([raw, k, k + 0.5, eighth, near, ulps])


def test_integer_expression_equals_the_double_sequence_when_not_ambiguous():
    rng = np.random.default_rng(20261017)
    raw = _cases(rng, 200_000)
    total_amb = 0
    for chroma in (False, True):
        for Q in (1, 2, 3, 8, 16, 31, 100, 255):
            P = rng.integers(-2300, 2301, raw.size)
            want = ref_level(raw, P, Q, chroma)
            got, amb = int_level(raw, P, Q, chroma)
            ok = ~amb
            assert np.array_equal(got[ok], want[ok]), f"chroma={chroma} Q={Q}: {np.count_nonzero(got[ok] != want[ok])} mismatches"
            total_amb += int(amb.sum())
    # the ambiguous set is exactly the constructed near-integer cases (integers - 0.5 ... + 2^-31), never a uniformly drawn value
    uniform = rng.uniform(-2041.0, 2041.0, 2_000_000)
    g = (uniform + 0.5) - np.floor(uniform + 0.5)
    assert np.count_nonzero((g < EPS) | (g > 1.0 - EPS)) <= 2       # expectation 2e6 * 2^-29 = 0.004
    assert total_amb > 0                                            # the test did exercise the ambiguous branch's mask


def test_ambiguous_margin_covers_the_rounding_of_the_double_sequence():
    """The flag must be set for every value whose double sequence COULD differ from the integer one: brute force over P for
    values within a few ulps of k + 0.5 shows mismatches only inside the flagged margin."""
    rng = np.random.default_rng(7)
    k = rng.integers(-2040, 2041, 20_000).astype(np.float64)
    parts, lo, hi = [k + 0.5], k + 0.5, k + 0.5
    for _ in range(6):                                              # up to 6 ulps to either side of k + 0.5
        lo, hi = np.nextafter(lo, -np.inf), np.nextafter(hi, np.inf)
        parts += [lo, hi]
    raw = np.concatenate(parts)
    for chroma in (False, True):
        P = rng.integers(-2300, 2301, raw.size)
        want = ref_level(raw, P, 1, chroma)
        got, amb = int_level(raw, P, 1, chroma)
        assert amb.all()                                            # all of these are inside the margin ...
        assert np.count_nonzero(got != want) > 0                    # ... and the margin is needed: some of them do differ
