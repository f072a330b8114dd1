"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's message-passing / pooling hot path
(Y-Claw/Multilevel-GNN), used as the parity checker.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import anything from here.  The product package
(``multilevel-gnn_b200/``) never imports it and has no CPU fallback.

Layout
------
``restated.py``     pure-torch CPU restatement of every op on the path; each
                    function cites the reference file:line it follows.  This is
                    what travels to the GPU box.
``pyg_stub.py``     container-only: a pure-torch stand-in for the third-party
                    packages the reference imports (torch_geometric 2.2.0,
                    torch_scatter 2.1.0, torch_cluster 1.6.0, h5py) so that the
                    reference's own files can be imported UNMODIFIED from
                    ``/root/reference``.
``ref_import.py``   container-only: imports reference modules behind the stub.
``make_golden.py``  container-only: runs the reference behind the stub on
                    seeded inputs and freezes inputs/outputs/grads under
                    ``tests/golden/``.

Pinning status: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4), and the arithmetic of torch_scatter / torch_geometric lives in
packages that are not vendored.  The restatement is therefore pinned against
OUTPUTS OF THE REFERENCE'S OWN FILES RUN HERE behind ``pyg_stub`` (the
committed ``tests/golden/*.pt`` + ``make_golden.py``), not against
reference-authored vectors.
"""
