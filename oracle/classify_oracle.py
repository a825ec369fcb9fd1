"""Oracle of the AUTO pre-classifier — TEST INFRASTRUCTURE ONLY.

numpy restatement of sdrpp_radiosonde_b200/csrc/classify.cu: run lengths between zero crossings of the DC-removed,
3-tap-smoothed discriminator output over the first <= 8192 samples, band fractions, decision rule.  There is no
reference counterpart (the reference's autodetect simply tries all seven decoders, SD/decode.c:174-224); what is
checked is (a) that the rule never excludes the true decoder type on the synthetic signals at any SNR and
(b) that the kernel computes this rule (tests/test_classifier.py)."""
import numpy as np

RS41, DFM09, M10, IMS100, MRZN1, IMET4, C50 = range(7)
ALL = 0x7F
NCLS = 8192


def classify_fm(fm):
    """fm: float32 discriminator output of one channel (one buffer).  Returns the 7-bit mask of plausible types."""
    m = min(len(fm), NCLS)
    if m < 2048:
        return ALL
    x = np.asarray(fm[:m], dtype=np.float32)
    mean = np.float32(x.sum(dtype=np.float32) / np.float32(m))
    d = np.concatenate([[np.float32(0)], x, [np.float32(0)]]).astype(np.float32)
    nw = (m // 32) * 32
    v = (d[0:nw] + d[1:nw + 1] + d[2:nw + 2]) * np.float32(1.0 / 3.0) - mean
    s = v > 0
    cross = np.nonzero(s[1:] != s[:-1])[0] + 1            # sample differs from its predecessor
    runs = np.diff(cross)
    total = len(runs)
    if total < 64:
        return ALL

    def frac(lo, hi):
        return np.count_nonzero((runs >= lo) & (runs <= hi)) / total

    r5, r8, r10, r20, r30 = frac(4, 6), frac(7, 9), frac(10, 12), frac(18, 22), frac(27, 33)
    if r5 >= 0.3:
        if r8 >= 0.2:
            return 1 << C50
        if r10 >= 0.15 and r8 < 0.1:
            return 1 << M10
        return ALL
    if r20 >= 0.45 and r10 < 0.1:
        return (1 << DFM09) | (1 << IMS100) | (1 << MRZN1)
    if r10 >= 0.25:
        if r30 >= 0.06:
            return 1 << RS41
        if r20 >= 0.1 and r30 < 0.02:
            return 1 << IMET4
        return (1 << RS41) | (1 << IMET4)
    return ALL
