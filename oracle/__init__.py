"""CPU oracle for the DSDGP ELBO / prediction path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package, and only as the checker / timed CPU baseline.  Nothing under
``dgplib_b200/`` imports it; the product path has no CPU fallback.

What it restates
----------------
The reference (aboustati/dgplib) is pure Python glue over GPflow (git master of mid-2018, about v1.2.0, unpinned:
``.travis.yml:15``) on TensorFlow 1.8.0 (``.travis.yml:14``).  Neither is vendored under ``/root/reference`` and
neither can be installed here (no network, no py3.12 build of TF 1.8).  The oracle therefore restates

* the dgplib half with file:line citations (``dgplib/dsdgp.py:84-130``, ``dgplib/layers.py:62-94``,
  ``dgplib/utils.py:15-27``, ``dgplib/multikernel_layers.py:65-110``, ``dgplib/specialized_kernels.py:23-72``,
  ``dgplib/multitask_dsdgp.py:9-26``, ``dgplib/layers.py:97-121``), and
* the GPflow half from its published algorithm (``conditional(..., white=True)``, ``gauss_kl``, ``RBF``,
  ``Matern52``, ``White``, ``Sum``, ``Gaussian``, ``SwitchedLikelihood``, ``Linear``/``Zero``), as written down in
  SURVEY.md Appendix A.

Pinning status
--------------
* PINNED against the reference's own exact goldens: the ``SwitchedKernel`` 10x5 gram matrix and ``Kdiag``
  (``tests/test_specialized_kernels.py:51-78``) and the three ``find_weights`` cases (``tests/test_layer.py:23-44``);
  see ``tests/test_oracle.py``.
* ELBO, gradients, predictive mean/variance: **parity unpinned** by the reference (its tests hold no numeric value
  for them; ``tests/test_layer.py:15-17`` is an empty stub, and TF/GPflow cannot run here to generate vectors).
  They are pinned instead by (i) the init-state identities that follow from ``dgplib/layers.py:52-56``
  (q_mu=0, q_sqrt=I => mean = mean_fn(X), var = Kdiag, KL = 0), (ii) the closed-form single-layer SVGP bound,
  (iii) an independent 50-digit mpmath evaluation on tiny cases (``oracle/mp_arbiter.py``, fixtures in
  ``tests/golden/``), (iv) finite differences for the autograd gradients, and (v) an independent third-party
  library: scikit-learn's RBF / Matern(nu=2.5) gram matrices (1e-12) and its exact-GP log marginal likelihood, which
  the single-layer ELBO must equal when Z = X and q(u) is the exact posterior (1e-8 relative), with its hyper-parameter
  gradients equal to scikit-learn's d lml / d theta and zero gradient in q (envelope theorem; ``tests/_exact_gp.py``).
"""
