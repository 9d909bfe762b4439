"""CPU oracle for the conditional RealNVP hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy / torch-CPU restatement of what the reference
(USArmyResearchLab/ARL_Conditional_Normalizing_Flows, TensorFlow/Keras) computes on
the path SURVEY.md §8 names: planner, masks, squeeze/factor, the ResNeXt s/t
sub-networks, the affine coupling law, the log-det, the flow and its loss, and the
toy dense model.  Every function cites the reference file:line it follows.

PARITY: PINNED TO THE REFERENCE'S SOURCE, UNPINNED AT TENSORFLOW'S KERNELS.  TensorFlow / TensorFlow-Probability / Keras
are not installable in this image and the reference ships no golden vectors, tests or saved weights, so this oracle cannot
be run against a real TensorFlow.  What it IS checked against: the reference's own model files
(/root/reference/conv_cINN_make_model.py, conv_cINN_base_functions.py), imported UNMODIFIED and executed in fp64 under the
NumPy stand-in of `oracle/tf_shim` (the ~35 tf ops and the Keras functional subset those files call, written from the
documented TensorFlow semantics).  `oracle/make_ref_golden.py` made `tests/golden/refsrc_*.npz` that way;
`tests/test_refsrc_golden.py` holds the restatement to them at 1e-12 (zy, the four loss scalars, log-det, samples) and the
CUDA path at 1e-4.  That pins the reference's CODE (planner, layer order, masks, squeeze / factor bookkeeping, net topology,
loss algebra, Keras variable order); TensorFlow's own conv / LayerNorm kernels remain documented-semantics re-implementations
until `tools/tf_dump_reference.py` is run on a TensorFlow box.  One open question found on the way -- whether tf.keras re-runs
the `Lambda` closure of F:402 with the loop variable's final value -- is spelled out in `oracle/tf_shim/README.md`; the
fixtures and `flow_torch.LAMBDA_LATE_BINDING` cover both readings.  Besides that the oracle is held by (i) the structural
invariants the reference's own code implies (tests/test_oracle_*), (ii) two independent restatements that must agree (a literal
NumPy transcription of the TF op sequence in `masks_np` / `nets_np`, and a torch functional version in `flow_torch`), and
(iii) autograd Jacobians on tiny shapes.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs may import this package.  The product
(`arl_conditional_normalizing_flows_b200`) never does.
"""
