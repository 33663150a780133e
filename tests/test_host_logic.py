"""Host-side behaviour mirrored from the reference's own unit tests (no GPU): find_weights goldens, initialize_forward,
Sequential validation rules, MultikernelLayer construction, DSDGP constructor checks, minibatch iterator."""
import numpy as np
import pytest
import torch

import dgplib_b200 as dg
from dgplib_b200.cascade import Sequential, MultitaskSequential
from dgplib_b200.kernels import RBF, Matern52, White
from dgplib_b200.layers import (find_weights, InputLayer, HiddenLayer, OutputLayer, Layer, SVGP_Layer)
from dgplib_b200.likelihoods import Gaussian, SwitchedLikelihood
from dgplib_b200.mean_functions import Linear, Zero
from dgplib_b200.multikernel_layers import (MultikernelLayer, MultikernelInputLayer, MultikernelHiddenLayer,
                                            MultikernelOutputLayer)
from dgplib_b200.specialized_kernels import SwitchedKernel


# ---- reference tests/test_layer.py
def test_find_weights():
    X = np.array([[2, 4], [1, 3], [0, 0], [0, 0]], dtype=float)
    W = find_weights(2, 2, X)
    assert W.shape == (2, 2) and np.allclose(W, np.eye(2))
    W = find_weights(2, 1, X)
    assert W.shape == (2, 1) and np.allclose(W, np.array([-0.4, -0.91]).reshape(2, 1), atol=1e-2)
    W = find_weights(1, 2, X)
    assert W.shape == (1, 2) and np.allclose(W, np.array([[1, 0]]))


def test_find_weights_matches_oracle_restatement():
    from oracle.dsdgp_oracle import find_weights as ofw
    rng = np.random.RandomState(0)
    X = rng.randn(30, 5)
    for (i, o, mt) in [(5, 5, False), (5, 2, False), (3, 5, False), (4, 2, True), (4, 4, True), (2, 4, True)]:
        Xi = X[:, :i + (1 if mt else 0)]
        assert np.allclose(find_weights(i, o, Xi, mt), ofw(i, o, Xi, mt))


@pytest.mark.parametrize("cls", [InputLayer, HiddenLayer])
def test_initialize_forward_input_hidden(cls):
    rng = np.random.RandomState(42)
    W0 = np.zeros((2, 2))
    Z, X = rng.randn(5, 2), rng.randn(10, 2)
    layer = cls(input_dim=2, output_dim=2, num_inducing=5, kernel=RBF(2), mean_function=Linear(A=W0))
    X_running, Z_running = layer.initialize_forward(X, Z)
    assert not np.allclose(layer.mean_function.A.value, W0)
    assert np.allclose(Z_running, Z) and np.allclose(layer.feature.Z.value, Z) and np.allclose(X_running, X)
    assert not layer.mean_function.A.trainable and not layer.mean_function.b.trainable   # layers.py:186


def test_initialize_forward_output():
    rng = np.random.RandomState(42)
    Z = rng.randn(5, 2)
    layer = OutputLayer(input_dim=2, output_dim=2, num_inducing=5, kernel=RBF(2))
    assert layer.initialize_forward(rng.randn(10, 2), Z) == (None, None)
    assert np.allclose(layer.feature.Z.value, Z)


def test_layer_parameters_at_construction():
    """dgplib/layers.py:36-56."""
    layer = Layer(3, 2, 7, RBF(3))
    assert SVGP_Layer is Layer
    assert layer.feature.Z.shape == (7, 3) and np.all(layer.feature.Z.value == 0)
    assert layer.q_mu.shape == (7, 2) and np.all(layer.q_mu.value == 0)
    assert layer.q_sqrt.shape == (2, 7, 7) and np.allclose(layer.q_sqrt.value, np.tile(np.eye(7), (2, 1, 1)))
    assert isinstance(layer.mean_function, Zero)
    assert Layer(3, 2, 7, RBF(3), multitask=True).feature.Z.shape == (7, 4)
    import torch
    if not torch.cuda.is_available():
        from dgplib_b200._cabi import DgpError
        with pytest.raises(DgpError):        # every compute path needs the GPU: no CPU fallback
            layer._build_predict(np.zeros((1, 2, 3)), full_cov=True)


# ---- reference tests/test_cascade.py
def _layers():
    return (InputLayer(2, 2, 10, RBF(2)), HiddenLayer(2, 2, 10, RBF(2)), OutputLayer(2, 1, 10, RBF(2)))


def test_sequential_validation_rules():
    il, hl, ol = _layers()
    with pytest.raises(AssertionError):
        Sequential([hl])                       # first layer must be an Input layer
    with pytest.raises(AssertionError):
        Sequential([il, "not a layer"])
    with pytest.raises(AssertionError):
        Sequential([il, HiddenLayer(3, 2, 10, RBF(3))])   # dims must chain
    seq = Sequential([il, hl, ol])
    with pytest.raises(ValueError):
        seq.add(HiddenLayer(1, 1, 10, RBF(1)))            # nothing after an Output layer
    assert seq.get_dims() == [(2, 2), (2, 2), (2, 1)]


def test_sequential_initialize_params():
    il, hl, ol = _layers()
    seq = Sequential([il, hl, ol])
    rng = np.random.RandomState(42)
    X, Z = rng.randn(20, 2), rng.randn(10, 2)
    seq.initialize_params(X, Z)
    assert seq._initialized
    for l in seq.layers:
        assert np.allclose(l.feature.Z.value, Z)
    with pytest.raises(ValueError):
        seq.add(OutputLayer(1, 1, 10, RBF(1)))            # cannot add to an initialised model


def test_multitask_sequential_keeps_id_column():
    rng = np.random.RandomState(42)
    X = np.hstack([rng.randn(20, 2), rng.randint(0, 2, (20, 1))])
    Z = np.hstack([rng.randn(10, 2), rng.randint(0, 2, (10, 1))])
    il = InputLayer(2, 1, 10, RBF(2), mean_function=Linear(A=np.ones((3, 1))), multitask=True)
    ol = OutputLayer(1, 1, 10, RBF(1), multitask=True)
    seq = MultitaskSequential([il, ol])
    seq.initialize_params(X, Z)
    for l in seq.layers:
        assert np.allclose(l.feature.Z.value[:, -1], Z[:, -1])
    assert il.feature.Z.shape == (10, 3) and ol.feature.Z.shape == (10, 2)
    assert il.mean_function.A.shape == (3, 1) and np.allclose(il.mean_function.A.value[-1], 0)


# ---- reference tests/test_multikernel_layer.py (construction / init parts)
def test_multikernel_layer_construction():
    with pytest.raises(ValueError):
        MultikernelLayer(2, 4, 5, [RBF(2), RBF(2), RBF(2)])       # 4 % 3 != 0
    shared = MultikernelInputLayer(2, 6, 5, [RBF(2), Matern52(2), RBF(2)], share_Z=True)
    assert shared.offset == 2 and shared.num_kernels == 3 and hasattr(shared.feature, "Z")
    sep = MultikernelInputLayer(2, 6, 5, [RBF(2), Matern52(2), RBF(2)], share_Z=False)
    assert len(sep.feature) == 3 and sep.feature[0].Z.shape == (5, 2)
    conds = sep.conditionals()
    assert [(c0, R) for (_, _, c0, R) in conds] == [(0, 2), (2, 2), (4, 2)]
    assert sep.q_mu.shape == (5, 6) and sep.q_sqrt.shape == (6, 5, 5)


@pytest.mark.parametrize("share", [True, False])
def test_multikernel_initialize_forward(share):
    rng = np.random.RandomState(42)
    X, Z = rng.randn(10, 2), rng.randn(5, 2)
    il = MultikernelInputLayer(2, 2, 5, [RBF(2), RBF(2)], share_Z=share, mean_function=Linear(A=np.zeros((2, 2))))
    hl = MultikernelHiddenLayer(2, 2, 5, [RBF(2), RBF(2)], share_Z=share)
    ol = MultikernelOutputLayer(2, 2, 5, [RBF(2), RBF(2)], share_Z=share)
    seq = Sequential([il, hl, ol])
    seq.initialize_params(X, Z)
    for l in (il, hl, ol):
        feats = [l.feature] if share else list(l.feature)
        for f in feats:
            assert np.allclose(f.Z.value, Z)
    assert np.allclose(il.mean_function.A.value, np.eye(2))


# ---- reference tests/test_dsdgp.py (constructor) and kernels / likelihood plumbing
def _toy():
    rng = np.random.RandomState(42)
    X = rng.uniform(0, 1, 50)[:, None]
    Z = rng.uniform(0, 1, 25)[:, None]
    Y = (X > 0.5).astype(float) + rng.randn(50, 1) * 1e-2
    return X, Y, Z


def test_dsdgp_constructor():
    X, Y, Z = _toy()
    seq = Sequential([InputLayer(1, 1, 25, RBF(1) + White(1)), OutputLayer(1, 1, 25, RBF(1) + White(1))])
    model = dg.DSDGP(X=X, Y=Y, Z=Z, layers=seq, likelihood=Gaussian())
    assert model.num_data == 50 and model.D_Y == 1 and model.dims == [(1, 1), (1, 1)]
    with pytest.raises(AssertionError):
        dg.DSDGP(X=X, Y=Y[:10], Z=Z, layers=Sequential([InputLayer(1, 1, 25, RBF(1))]), likelihood=Gaussian())
    with pytest.raises(AssertionError):
        dg.DSDGP(X=X, Y=Y, Z=np.zeros((25, 2)), layers=Sequential([InputLayer(1, 1, 25, RBF(1))]), likelihood=Gaussian())


def test_minibatch_iterator_keeps_x_and_y_aligned():
    X, Y, Z = _toy()
    Y = np.hstack([Y, X])      # second column of Y mirrors X so that alignment can be checked
    seq = Sequential([InputLayer(1, 1, 25, RBF(1)), OutputLayer(1, 1, 25, RBF(1))])
    model = dg.DSDGP(X=X, Y=Y, Z=Z, layers=seq, likelihood=Gaussian(), minibatch_size=16, num_latent=1)
    seen = []
    for _ in range(7):          # crosses an epoch boundary (50 rows / 16)
        xb, yb = model._next_batch()
        assert xb.shape == (16, 1) and np.allclose(xb[:, 0].numpy(), yb[:, 1].numpy())
        seen.append(xb[:, 0].numpy())
    first_epoch = np.concatenate(seen)[:50]
    assert len(np.unique(first_epoch)) == 50       # a permutation: every row once per epoch


def test_positive_transform_and_kernel_parts():
    k = RBF(3, variance=2.0, lengthscales=[1.0, 2.0, 3.0]) + White(3, variance=1e-5)
    kind, var, white, ls = k.stationary_parts()
    assert kind == 0 and abs(float(var) - 2.0) < 1e-12 and abs(float(white) - 1e-5) < 1e-15
    assert np.allclose(ls.detach().numpy(), [1, 2, 3])
    k.kernels[0].variance.assign(3.5)
    assert abs(float(k.kernels[0].variance.value) - 3.5) < 1e-12
    with pytest.raises(NotImplementedError):
        (RBF(2) + RBF(2)).stationary_parts()
    sw = SwitchedKernel([RBF(2), Matern52(2)], 2)
    assert sw.switched and [p[0] for p in sw.task_kernels()] == [0, 1]
    with pytest.raises(AssertionError):
        SwitchedKernel([RBF(2), RBF(2)], 3)
    with pytest.raises(NotImplementedError):
        SwitchedLikelihood([Gaussian(), object()])


def test_compute_paths_fail_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from dgplib_b200._cabi import DgpError
    X, Y, Z = _toy()
    seq = Sequential([InputLayer(1, 1, 25, RBF(1)), OutputLayer(1, 1, 25, RBF(1))])
    model = dg.DSDGP(X=X, Y=Y, Z=Z, layers=seq, likelihood=Gaussian())
    for call in (lambda: model.compute_log_likelihood(), lambda: model.predict_f(X, 1), lambda: model.predict_y(X),
                 lambda: model.predict_f_full_cov(X, 1), lambda: model.predict_f_samples(X, 3),
                 lambda: model.elbo_and_grad()):
        with pytest.raises(DgpError):
            call()


def test_structure_epoch_tracks_every_change_of_the_parameter_set():
    """The engine caches its walk over the module tree on params.structure_epoch(): every way the SET of torch Parameters
    of a model can change has to bump it (in-place value updates must not)."""
    from dgplib_b200 import params as P
    from dgplib_b200.kernels import RBF, White

    def bumped(fn):
        e0 = P.structure_epoch()
        fn()
        return P.structure_epoch() > e0

    k = RBF(2)
    assert bumped(lambda: RBF(2))                                         # new Params
    assert not bumped(lambda: k.variance.assign(2.0))                     # in place
    assert bumped(lambda: k.lengthscales.assign(np.ones(2)))              # shape change: `raw` is replaced
    assert bumped(lambda: setattr(k, "variance", P.Param(1.5, positive=True)))   # a sub-module replaced
    assert bumped(lambda: setattr(k.variance, "raw", torch.nn.Parameter(torch.zeros((), dtype=torch.float64))))
    s = k + White(2)
    assert bumped(lambda: s.kernels.append(White(2)))                     # ParamList edited
    assert bumped(lambda: s.kernels.__setitem__(0, RBF(2)))
    layer = InputLayer(2, 2, 5, RBF(2))
    assert bumped(lambda: setattr(layer, "kernel", RBF(2)))
    assert not bumped(lambda: layer.q_mu.assign(np.ones((5, 2))))
