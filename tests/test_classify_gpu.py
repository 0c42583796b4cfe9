"""classify pre-processing (kameris/job_steps/classify.py:29-50) on the device vs scikit-learn itself: the
scaler's statistics / transform and the randomized TruncatedSVD for the same random_state.  scikit-learn is
the reference's own dependency for this arithmetic (setup.py:85), so the installed copy is the oracle."""
import io
import pickle

import numpy as np
import pytest

from oracle import oracle as O
from tests.helpers import HIV_P, synth_records

torch = pytest.importorskip("torch")
sk_pre = pytest.importorskip("sklearn.preprocessing")
sk_dec = pytest.importorskip("sklearn.decomposition")


def _counts(n, L, k, seed, bits=16):
    recs = synth_records(n, L, seed=seed, p=HIV_P, jitter=L // 10)
    return np.stack([O.cgr_count(r, k, bits) for r in recs])


# ---- host logic (no GPU) ------------------------------------------------------------------------------
def test_build_pipeline_mirrors_reference_structure():
    from sklearn.svm import LinearSVC
    from kameris_b200.job_steps import classify as C
    p = C.build_pipeline(LinearSVC, 3521, {})
    assert [n for n, _ in p.steps] == ["scaler", "dim_reducer", "classifier"]
    assert p.named_steps["dim_reducer"].n_components == 353            # ceil(3521 * 0.1), classify.py:45-47
    assert p.named_steps["scaler"].with_mean is False
    p = C.build_pipeline(LinearSVC, 100, {"dim_reduce_fraction": 0.25})
    assert p.named_steps["dim_reducer"].n_components == 25
    p = C.build_pipeline(LinearSVC, 100, {"skip_normalization": True})
    assert [n for n, _ in p.steps] == ["classifier"]
    feats = [np.array([0, 1, 2, 0]), np.array([1, 1, 1, 1]), np.array([0, 0, 0, 5])]
    assert C.avg_num_nonzero_entries(feats) == 2                       # int(7 / 3), classify.py:25-26


def test_random_state_matches_sklearn():
    from sklearn.utils import check_random_state
    from kameris_b200 import preprocess as P
    a = P.random_state_of(11).normal(size=(5, 3))
    b = check_random_state(11).normal(size=(5, 3))
    assert np.array_equal(a, b)
    rs = np.random.RandomState(3)
    assert P.random_state_of(rs) is rs and P.random_state_of(None) is check_random_state(None)


def test_no_cpu_fallback():
    from kameris_b200 import preprocess as P
    with pytest.raises(RuntimeError):
        P.column_stats(torch.zeros((4, 4), dtype=torch.float64))
    with pytest.raises(RuntimeError):
        P.randomized_svd(torch.zeros((4, 4), dtype=torch.float64), 2)


# ---- device vs scikit-learn ---------------------------------------------------------------------------
needs_gpu = pytest.mark.gpu


@pytest.fixture(scope="module")
def C():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from kameris_b200.job_steps import classify as C
    return C


@needs_gpu
@pytest.mark.parametrize("case", ["u16_k6", "u32_k4", "f64_freq_k5", "f32_freq_k5", "f64_nan_const", "tall_thin", "one_row",
                                  "u8_vec4", "u8_odd"])
def test_standard_scaler_vs_sklearn(C, case):
    rng = np.random.Generator(np.random.PCG64(5))
    if case == "u16_k6":
        x = _counts(90, 9000, 6, seed=1)
    elif case == "u32_k4":
        x = _counts(300, 3000, 4, seed=2, bits=32)
    elif case in ("f64_freq_k5", "f32_freq_k5"):
        c = _counts(150, 5000, 5, seed=3)
        x = (c / c.sum(axis=1, keepdims=True)).astype(np.float64 if case.startswith("f64") else np.float32)
    elif case == "f64_nan_const":
        x = rng.standard_normal((70, 37)) * 3 + 10
        x[::7, 3] = np.nan                         # missing values are skipped per column
        x[:, 5] = 2.5                              # constant column: scale 1
        x[:, 6] = 0.0
        x[:, 8] = 1e9 + rng.standard_normal(70) * 1e-9      # indistinguishable from a constant: scale 1
    elif case in ("u8_vec4", "u8_odd"):            # 4 columns per thread / the one-column fallback
        x = rng.integers(0, 200, size=(777, 1028 if case == "u8_vec4" else 1027)).astype(np.uint8)
    elif case == "tall_thin":
        x = rng.integers(0, 50, size=(5000, 3)).astype(np.uint16)
    else:
        x = rng.integers(0, 9, size=(1, 300)).astype(np.uint16)
    ref = sk_pre.StandardScaler(with_mean=False).fit(x)
    got = C.StandardScaler(with_mean=False).fit(x)
    np.testing.assert_allclose(got.mean_, ref.mean_, rtol=1e-13, atol=0)
    # two-pass variance: relative agreement down to the rounding floor n * eps * mean^2 of the formula itself
    floor = 1e-15 * x.shape[0] * np.nan_to_num(ref.mean_) ** 2
    assert np.all(np.abs(got.var_ - ref.var_) <= 1e-10 * np.abs(ref.var_) + floor)
    assert np.array_equal(got.scale_ == 1.0, ref.scale_ == 1.0)
    np.testing.assert_allclose(got.scale_, ref.scale_, rtol=1e-10)
    assert np.array_equal(np.atleast_1d(got.n_samples_seen_), np.atleast_1d(ref.n_samples_seen_))
    t_ref = ref.transform(x.astype(np.float64) if x.dtype.kind == "u" else x.copy())
    t_got = got.transform(x)
    assert t_got.dtype == np.float64 and t_got.shape == x.shape
    # (sklearn keeps a float32 input in float32: its own transform is only good to 6e-8 there)
    np.testing.assert_allclose(t_got, t_ref, rtol=1e-10 if x.dtype != np.float32 else 3e-7, equal_nan=True)


@needs_gpu
@pytest.mark.parametrize("n,d,comp,seed", [(120, 1024, 30, 0), (700, 256, 25, 1), (64, 4096, 50, 2), (400, 400, 10, 3)])
def test_truncated_svd_vs_sklearn(C, n, d, comp, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    base = rng.poisson(2.0, size=(n, d)).astype(np.float64)
    base[: n // 2] += rng.poisson(1.0, size=(1, d))                   # two groups: a spectrum with structure
    x = sk_pre.StandardScaler(with_mean=False).fit_transform(base)
    ref = sk_dec.TruncatedSVD(n_components=comp, n_iter=5, random_state=seed)
    t_ref = ref.fit_transform(x)
    got = C.TruncatedSVD(n_components=comp, n_iter=5, random_state=seed)
    t_got = got.fit_transform(x)
    np.testing.assert_allclose(got.singular_values_, ref.singular_values_, rtol=1e-9)
    scale = np.abs(ref.components_).max()
    np.testing.assert_allclose(got.components_, ref.components_, atol=1e-8 * scale, rtol=0)
    np.testing.assert_allclose(t_got, t_ref, atol=1e-8 * np.abs(t_ref).max(), rtol=0)
    np.testing.assert_allclose(got.explained_variance_ratio_, ref.explained_variance_ratio_, rtol=1e-7)
    y = x[: n // 3] * 1.5
    np.testing.assert_allclose(got.transform(y), ref.transform(y), atol=1e-8 * np.abs(t_ref).max(), rtol=0)
    with pytest.raises(ValueError):
        C.TruncatedSVD(n_components=d + 1).fit(x)


@needs_gpu
def test_pipeline_predictions_equal_sklearn_pipeline(C):
    """build_pipeline (device transformers) against the reference's pipeline built from sklearn's own
    transformers (classify.py:29-50), same random_state: same reduced features, same predictions; the fitted
    pipeline survives the pickling of classify.py:254."""
    from sklearn.neighbors import KNeighborsClassifier
    from sklearn.pipeline import Pipeline
    a = _counts(60, 9000, 5, seed=10)
    recs_b = synth_records(60, 9000, seed=11, p=[0.25, 0.25, 0.25, 0.25], jitter=900)
    b = np.stack([O.cgr_count(r, 5, 16) for r in recs_b])
    x = np.concatenate([a, b])
    y = np.array(["A"] * 60 + ["B"] * 60)
    perm = np.random.Generator(np.random.PCG64(0)).permutation(120)
    tr, te = perm[:90], perm[90:]
    feats = [x[i] for i in range(120)]                                  # classify.py:191-196: a list of vectors
    num = C.avg_num_nonzero_entries(feats)
    ours = C.build_pipeline(lambda: KNeighborsClassifier(3), num, {})
    ours.set_params(dim_reducer__random_state=4)
    ref = Pipeline([("scaler", sk_pre.StandardScaler(with_mean=False)),
                    ("dim_reducer", sk_dec.TruncatedSVD(n_components=int(np.ceil(num * 0.1)), random_state=4)),
                    ("classifier", KNeighborsClassifier(3))])
    ours.fit([feats[i] for i in tr], y[tr])
    ref.fit(np.stack([feats[i] for i in tr]).astype(np.float64), y[tr])
    test = [feats[i] for i in te]
    p_ours, p_ref = ours.predict(test), ref.predict(np.stack(test).astype(np.float64))
    assert np.array_equal(p_ours, p_ref) and (p_ours == y[te]).mean() > 0.9
    np.testing.assert_allclose(ours.predict_proba(test), ref.predict_proba(np.stack(test).astype(np.float64)), atol=1e-9)
    np.testing.assert_allclose(np.sum(ours.named_steps["dim_reducer"].explained_variance_ratio_),
                               np.sum(ref.named_steps["dim_reducer"].explained_variance_ratio_), rtol=1e-7)
    again = pickle.load(io.BytesIO(pickle.dumps(ours)))
    assert np.array_equal(again.predict(test), p_ours)


@needs_gpu
def test_scaler_full_size_property(C):
    """10 000 x 4096 uint16 count vectors: unit variance of every non-constant column after the transform
    (the defining property), column means = mean / scale."""
    from kameris_b200 import preprocess as P
    g = torch.Generator(device="cuda")
    g.manual_seed(3)
    x = torch.randint(0, 12, (10000, 4096), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
    mean, var, cnt = P.column_stats(x)
    scale = P.scale_from_variance(var, mean, cnt)
    t = P.scale_columns(x, scale)
    assert int(cnt.min()) == 10000 and int(cnt.max()) == 10000
    torch.testing.assert_close(t.var(dim=0, unbiased=False), torch.ones(4096, dtype=torch.float64, device="cuda"),
                               rtol=1e-10, atol=0)
    torch.testing.assert_close(t.mean(dim=0), mean / scale, rtol=1e-10, atol=0)


@needs_gpu
@pytest.mark.parametrize("dtype", ["u16", "f64", "f32"])
def test_scale_columns_is_a_correctly_rounded_division(C, dtype):
    """kam_col_scale forms x / scale from a reciprocal and one exact-remainder correction: the results must be
    the IEEE quotients bit for bit (checked against the device's own division), including zeros, NaN,
    infinities, denormal and huge operands and scales."""
    from kameris_b200 import preprocess as P
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    n, d = 3000, 515
    if dtype == "u16":
        x = torch.randint(0, 65536, (n, d), device="cuda", generator=g, dtype=torch.int32).to(torch.uint16)
        xd = x.to(torch.int32).to(torch.float64)
    else:
        x = torch.randn((n, d), device="cuda", generator=g, dtype=torch.float64) * 10 ** torch.randint(
            -12, 12, (n, d), device="cuda", generator=g).to(torch.float64)
        x[5, :] = 0.0
        x[6, 3] = float("nan")
        x[7, 4] = float("inf")
        x[8, 5] = 1e-310
        x[9, 6] = 1e300
        if dtype == "f32":
            x = x.to(torch.float32)
        xd = x.to(torch.float64)
    scale = torch.rand(d, device="cuda", generator=g, dtype=torch.float64) * 10 ** torch.randint(
        -6, 6, (d,), device="cuda", generator=g).to(torch.float64) + 1e-9
    scale[0], scale[1], scale[2] = 1.0, 1e-200, 1e200
    got = P.scale_columns(x, scale)
    want = xd / scale
    assert torch.equal(torch.nan_to_num(got, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
    assert torch.equal(torch.isnan(got), torch.isnan(want))
