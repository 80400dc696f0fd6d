"""CPU oracle for the TensorFlowASR training hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``tensorflowasr_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and only as the checker / the CPU
arm, never as the product path.

Each module restates, in numpy / torch-CPU / plain C, the arithmetic of the
reference files it cites (paths relative to ``/root/reference``).  TensorFlow and
Keras are not installable in the build container, so the reference itself cannot
be executed: every module header states what its parity is pinned against
(reference golden vectors where they exist, independent third-party
implementations otherwise) or says "parity unpinned".
"""
