"""Drop-in for the pre-processing half of kameris/job_steps/classify.py: the pipeline that
`build_pipeline` (classify.py:29-50) puts in front of every classifier,

    StandardScaler(with_mean=False)  ->  TruncatedSVD(n_components=ceil(num_features * dim_reduce_fraction))

as scikit-learn-compatible transformers whose arithmetic runs on the B200 (kameris_b200/preprocess.py):
same constructor arguments, same fitted attributes (mean_, var_, scale_, n_samples_seen_; components_,
singular_values_, explained_variance_, explained_variance_ratio_), same results as the installed
scikit-learn for the same `random_state` (tests/test_classify_gpu.py).  The classifiers themselves, the
cross-validation loop and the model files stay with kameris / scikit-learn: `build_pipeline` here has the
reference's signature and returns an sklearn Pipeline with these two transformers in place of sklearn's.

Host arrays (the lists of count / frequency vectors read from a `.mm-repr`, classify.py:191-196) are
uploaded in their stored element type; between the two transformers the matrix stays on the device.
"""
from __future__ import absolute_import, division, unicode_literals

import numpy as np
from sklearn.base import BaseEstimator, TransformerMixin
from sklearn.pipeline import Pipeline


def avg_num_nonzero_entries(features):
    """classify.py:25-26."""
    return int(sum(np.count_nonzero(f) for f in features) / len(features))


def _to_device(X):
    """host matrix / list of vectors / CUDA tensor -> CUDA tensor in the stored element type."""
    import torch
    if isinstance(X, torch.Tensor):
        if not X.is_cuda:
            X = X.cuda()
        return X
    a = np.asarray(X)
    if a.ndim != 2:
        raise ValueError("Expected 2D array, got %dD array instead" % a.ndim)
    if a.dtype == np.uint64 or a.dtype == np.int64:
        a = a.astype(np.float64)
    elif a.dtype.kind not in "uif":
        raise ValueError("could not convert the feature vectors to numbers")
    a = np.ascontiguousarray(a)
    view = {np.dtype(np.uint16): (np.int16, torch.uint16), np.dtype(np.uint32): (np.int32, torch.uint32)}
    if a.dtype in view:
        return torch.from_numpy(a.view(view[a.dtype][0])).cuda().view(view[a.dtype][1])
    if a.dtype == np.float16:
        a = a.astype(np.float32)
    return torch.from_numpy(a).cuda()


class StandardScaler(TransformerMixin, BaseEstimator):
    """sklearn.preprocessing.StandardScaler(with_mean=False, with_std=True) on the device.
    `device_output=True` makes transform() return the CUDA tensor (for a following device transformer)."""

    def __init__(self, copy=True, with_mean=False, with_std=True, device_output=False):
        self.copy = copy
        self.with_mean = with_mean
        self.with_std = with_std
        self.device_output = device_output

    def fit(self, X, y=None):
        from .. import preprocess as P
        if self.with_mean or not self.with_std:
            raise ValueError("only with_mean=False, with_std=True is implemented (what kameris uses: "
                             "classify.py:40)")
        x = _to_device(X)
        mean, var, cnt = P.column_stats(x)
        self.scale_device_ = P.scale_from_variance(var, mean, cnt)
        self.mean_ = mean.cpu().numpy()
        self.var_ = var.cpu().numpy()
        self.scale_ = self.scale_device_.cpu().numpy()
        cnt = cnt.cpu().numpy()
        self.n_samples_seen_ = int(cnt[0]) if cnt.min() == cnt.max() else cnt
        self.n_features_in_ = x.shape[1]
        return self

    def transform(self, X, copy=None):
        from .. import preprocess as P
        x = _to_device(X)
        if x.shape[1] != self.n_features_in_:
            raise ValueError("X has %d features, but StandardScaler is expecting %d features as input."
                             % (x.shape[1], self.n_features_in_))
        scale = getattr(self, "scale_device_", None)
        if scale is None or scale.device != x.device:      # unpickled model, or another GPU: upload once
            import torch
            scale = self.scale_device_ = torch.from_numpy(np.asarray(self.scale_, dtype=np.float64)).to(x.device)
        out = P.scale_columns(x, scale)
        return out if self.device_output else out.cpu().numpy()

    def __getstate__(self):                      # model files (.mm-model, classify.py:254) hold host arrays only
        state = dict(self.__dict__)
        state.pop("scale_device_", None)
        return state



class TruncatedSVD(TransformerMixin, BaseEstimator):
    """sklearn.decomposition.TruncatedSVD(algorithm='randomized') on the device."""

    def __init__(self, n_components=2, algorithm="randomized", n_iter=5, n_oversamples=10, random_state=None, tol=0.0):
        self.n_components = n_components
        self.algorithm = algorithm
        self.n_iter = n_iter
        self.n_oversamples = n_oversamples
        self.random_state = random_state
        self.tol = tol

    def _as_f64(self, X):
        import torch
        x = _to_device(X)
        return x if x.dtype == torch.float64 else x.to(torch.float64)

    def fit(self, X, y=None):
        self.fit_transform(X)
        return self

    def fit_transform(self, X, y=None):
        from .. import preprocess as P
        if self.algorithm != "randomized":
            raise ValueError("only algorithm='randomized' (sklearn's default, what kameris uses) is implemented")
        x = self._as_f64(X)
        r = P.truncated_svd_fit(x, self.n_components, self.n_iter, self.n_oversamples, self.random_state)
        self.components_device_ = r["components_"]
        self.components_ = r["components_"].cpu().numpy()
        self.singular_values_ = r["singular_values_"].cpu().numpy()
        self.explained_variance_ = r["explained_variance_"].cpu().numpy()
        self.explained_variance_ratio_ = r["explained_variance_ratio_"].cpu().numpy()
        self.n_features_in_ = x.shape[1]
        return r["transformed"].cpu().numpy()

    def transform(self, X):
        x = self._as_f64(X)
        if x.shape[1] != self.n_features_in_:
            raise ValueError("X has %d features, but TruncatedSVD is expecting %d features as input."
                             % (x.shape[1], self.n_features_in_))
        comp = getattr(self, "components_device_", None)
        if comp is None or comp.device != x.device:          # unpickled model, or another GPU: upload once
            import torch
            comp = self.components_device_ = torch.from_numpy(
                np.asarray(self.components_, dtype=np.float64)).to(x.device)
        return (x @ comp.T).cpu().numpy()

    def inverse_transform(self, X):
        return np.dot(X, self.components_)

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("components_device_", None)
        return state



def build_pipeline(classifier_factory, num_features, options):
    """Same contract as classify.py:29-50: options `skip_normalization` (default False) and
    `dim_reduce_fraction` (default 0.1); unless normalisation is skipped, the classifier is preceded by the
    unit-variance scaler (no centring) and by a truncated SVD keeping ceil(num_features * fraction)
    components, here both on the device with the scaled matrix handed over without leaving it."""
    steps = []
    if not options.get('skip_normalization', False):
        keep = int(np.ceil(num_features * options.get('dim_reduce_fraction', 0.1)))
        steps += [('scaler', StandardScaler(with_mean=False, device_output=True)),
                  ('dim_reducer', TruncatedSVD(n_components=keep))]
    steps.append(('classifier', classifier_factory()))
    return Pipeline(steps)
