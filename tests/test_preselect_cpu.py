"""CPU: the top-max_nms pre-selection of the filter kernels (csrc/nms.cu: key_bin, hist_threshold_kernel, bin test) restated in
numpy -- the bin function is monotone in the sort key, and thresholding at the bin of the K-th best key keeps a superset of
the top K (ties included) in candidate order, so the stable sort + truncation downstream returns the same candidates."""
import numpy as np
import pytest

KBINS = 2048


def key_bin(conf):
    """csrc/nms.cu key_bin(): top 15 bits of ~bits(conf) minus those of 1.0, clamped to [0, 2047]."""
    bits = np.atleast_1d(np.asarray(conf, dtype=np.float32)).view(np.uint32)
    k = ((~bits) >> np.uint32(17)).astype(np.int64)
    base = int((~np.uint32(0x3F800000)) >> np.uint32(17))
    out = np.clip(k - base, 0, KBINS - 1)
    return out if np.ndim(conf) else int(out[0])


def bin_threshold(bins, K):
    """hist_threshold_kernel: first bin whose inclusive count reaches K (the last bin when there are at most K candidates)."""
    inc = np.cumsum(np.bincount(bins, minlength=KBINS))
    hit = np.nonzero(inc >= K)[0]
    return int(hit[0]) if len(hit) else KBINS - 1


def test_key_bin_is_monotone_in_the_sort_key():
    rng = np.random.default_rng(0)
    conf = np.concatenate([rng.random(20000, dtype=np.float32), np.float32(2.0) ** -rng.integers(0, 40, 2000),
                           np.array([1.0, 0.999999, 0.5, 0.25, 1e-3, 1e-6, 1e-12, 1e-30, 1e-45, 1.5, 3.0, 100.0], np.float32)])
    conf = conf[conf > 0]
    order = np.argsort(-conf.astype(np.float64), kind='stable')          # best score first == ascending sort key
    b = key_bin(conf[order])
    assert np.all(np.diff(b) >= 0)                                        # better score -> lower or equal bin
    assert key_bin(np.float32(1.0)) == 0 and key_bin(np.float32(7.0)) == 0 and key_bin(np.float32(1e-30)) == KBINS - 1
    one_binade = key_bin(np.float32(0.5) + np.arange(64, dtype=np.float32) / 128)   # 64 bins per binade
    assert len(np.unique(one_binade)) == 64


@pytest.mark.parametrize('n,K,levels', [(5000, 300, 0), (5000, 300, 7), (200, 300, 0), (40000, 30000, 0), (40000, 30000, 50)])
def test_bin_threshold_keeps_a_superset_of_the_top_k_in_order(n, K, levels):
    rng = np.random.default_rng(n + K + levels)
    conf = (rng.random(n, dtype=np.float32) * rng.random(n, dtype=np.float32)).astype(np.float32)
    if levels:                                                            # massive ties
        conf = (np.ceil(conf * levels) / levels).astype(np.float32)
    conf = conf[conf > 1e-3]
    bins = key_bin(conf)
    keep = bins <= bin_threshold(bins, K)
    ref = np.argsort(-conf.astype(np.float64), kind='stable')[:K]         # utils/general.py:702-703 with a stable order
    kept_idx = np.nonzero(keep)[0]
    got = kept_idx[np.argsort(-conf[kept_idx].astype(np.float64), kind='stable')[:K]]
    assert np.array_equal(got, ref)
    assert keep.sum() >= min(K, len(conf))
