"""CPU oracle for the Fast-SCNN / ContextNet hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product
package ``torch_semantic_segmentation_b200``; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it, and there only as the checker / the CPU arm.

The oracle is a *functional restatement* (plain ``torch.nn.functional`` fp32 on CPU,
keyed by the reference's ``state_dict`` names) of:

* ``torch_semantic_segmentation/models/fastscnn.py``   -> :mod:`oracle.fastscnn`
* ``torch_semantic_segmentation/models/contextnet.py`` -> :mod:`oracle.contextnet`
* ``torch.nn.CrossEntropyLoss(ignore_index=255)`` (``scripts/train_fastscnn.py:132``)
  and ``losses/ohem_loss.py``                           -> :mod:`oracle.losses`
* ignite ``ConfusionMatrix/IoU/mIoU/cmAccuracy/DiceCoefficient`` used at
  ``engine.py:65-72`` (third party, NOT in /root/reference, un-pinned)
                                                        -> :mod:`oracle.confusion`
* ``engine.py:24-39`` ``update_fn``                     -> :mod:`oracle.train_step`

Pinning: ``oracle/make_golden.py`` imports the UNMODIFIED reference from
``/root/reference`` (possible only in the build container), runs it on seeded inputs
and stores its outputs under ``tests/golden/``; ``tests/test_oracle.py`` checks the
restatement against those vectors everywhere, and against the live reference when
``/root/reference`` is present.  The confusion-matrix part has no reference source to
run (ignite is absent and un-pinned) -> that part is "parity unpinned" by the
reference; it is cross-checked against scikit-learn's independent implementation.
"""
