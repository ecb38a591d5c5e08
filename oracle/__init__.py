"""CPU oracle for the learning-to-rank hot path.  TEST INFRASTRUCTURE ONLY.

This package is a torch-CPU restatement (fp32 or fp64, chosen by the dtype of
the tensors passed in) of the reference algorithm in tensorflow/ranking for the
path named by BASELINE.json `north_star`:

    scorer MLP -> pairwise / ApproxNDCG / Softmax loss -> NDCG / MRR metrics

It materialises the same [B, N, N] pairwise tensors the reference does
(`losses_impl.py:61-64`) and relies on torch autograd for gradients, exactly as
the reference relies on TF autodiff.  Each function cites the reference
file:line it follows (paths relative to tensorflow_ranking/python/).

Who may import it: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and only as the checker or the timed CPU
baseline.  Nothing under `ranking_b200/` imports it; the product path raises if
its CUDA library is missing instead of falling back to this code.

Parity pinning: TensorFlow is not installable in the authoring container, so the
oracle is pinned against the reference's OWN tests, whose expected values are
closed-form Python `math` expressions independent of TF (`losses_impl_test.py`,
`keras/losses_test.py`, `metrics_impl_test.py`, `utils_test.py`,
`keras/layers_test.py`, `model_test.py`).  Those vectors are ported in
`tests/test_oracle_*.py` and frozen in `tests/golden/reference_vectors.json`.
What stays unpinned (the reference has no portable test for it): gradients
(only autograd-vs-kernel is checked, plus sign golden of `model_test.py:392-399`)
and random tie-breaking (TF Philox stream); ties are broken index-stable here,
i.e. the reference's `shuffle_ties=False` semantics (`utils_test.py:104-126`).
"""
