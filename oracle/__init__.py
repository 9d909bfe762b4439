"""CPU oracle for the conditional RealNVP hot path — TEST INFRASTRUCTURE ONLY.

This package is a plain NumPy / torch-CPU restatement of what the reference
(USArmyResearchLab/ARL_Conditional_Normalizing_Flows, TensorFlow/Keras) computes on
the path SURVEY.md §8 names: planner, masks, squeeze/factor, the ResNeXt s/t
sub-networks, the affine coupling law, the log-det, the flow and its loss, and the
toy dense model.  Every function cites the reference file:line it follows.

PARITY UNPINNED: TensorFlow / TensorFlow-Probability / Keras are not installable in
this image and the reference ships no golden vectors, tests or saved weights, so
this oracle cannot be run against the reference itself.  It is pinned instead by
(i) the structural invariants the reference's own code implies (tests/test_oracle_*),
(ii) two independent restatements that must agree (a literal NumPy transcription of
the TF op sequence in `masks_np` / `nets_np`, and a torch functional version in
`flow_torch`), and (iii) autograd Jacobians on tiny shapes.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / --impl
reference legs may import this package.  The product
(`arl_conditional_normalizing_flows_b200`) never does.
"""
