"""CPU oracle for the semantic-embeddings hot path.  TEST INFRASTRUCTURE ONLY.

This package is a float64 (optionally float32) CPU restatement of what the
reference (cvjena/semantic-embeddings, Keras 2.2 on TensorFlow 1.x) computes on
the path named by BASELINE.json: CNN forward/backward -> L2-normalise ->
1-cosine loss against the fixed class-embedding matrix -> clipped momentum SGD,
and the all-pairs distance matrix + ranking of evaluate_retrieval.py.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` leg may import it -- as the checker (or the timed CPU
baseline), never as part of the product path.  Nothing in
`semantic_embeddings_b200/` imports it.

Parity pinning status (see DESIGN.md "Oracle"):
  * retrieval (distance + argsort ranking + hierarchical precision): PINNED
    against the reference's own `evaluate_retrieval.pairwise_retrieval` and
    `class_hierarchy.ClassHierarchy` executed in the build container
    (tests/golden/make_golden.py, fixtures under tests/golden/).
  * SGDR schedule, loss / metric formulas (`utils.inv_correlation`,
    `utils.nn_accuracy`), architecture layer lists: PINNED against the
    reference's own Python executed under a recording stub of the (absent)
    Keras API (tests/golden/make_golden.py).
  * Numerical semantics of the un-vendored Keras/TensorFlow layers (Conv2D SAME
    padding, BatchNormalization, SGD(clipnorm)): restated from the published
    behaviour of keras==2.2 / tensorflow 1.x; the reference holds no golden
    vectors for them and Keras/TF cannot be installed here => "parity unpinned"
    for those layer semantics.
"""
