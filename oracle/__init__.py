"""Oracle for the GripNet supergraph message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  ``gripnet_b200`` never imports this package.

Contents
--------
``pyg_shim``    three third-party symbols the reference imports (PyG 1.x
                ``MessagePassing`` / ``add_remaining_self_loops`` and
                ``torch_scatter.scatter_add``), restated from their published
                behaviour so the UNMODIFIED reference can be imported from
                ``/root/reference`` in the build container (``ref_loader``).
``port``        CPU restatement (torch-CPU / numpy) of every function on the
                hot path, each citing the reference file:line it follows.
                This is what travels to the GPU box.
``dense64``     independent float64 dense-matrix restatement (numpy), used to
                arbitrate tolerance questions.
``synth``       seeded synthetic supergraph generators for the BASELINE configs.

Parity pin status: the reference ships no tests or golden vectors
 and its arithmetic lives in un-vendored, un-pinned PyG<2.0 /
torch_scatter.  The pin used here is: outputs of the reference's own
``gripnet/*.py`` run in the build container over ``pyg_shim`` and committed as
``tests/golden/*.npz`` by ``tests/golden/make_golden.py``; ``port`` and
``dense64`` are both checked against those fixtures.  The shim itself is a
restatement of third-party semantics, so at the PyG boundary parity is
"pinned to the reference source + restated PyG-1.x semantics", not to the
real PyG binaries (which cannot be installed here).
"""
