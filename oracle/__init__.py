"""
oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the sampling hot path of unixpickle/vq-voice-swap (SURVEY.md
section 8a).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import this
package; the product (``vq_voice_swap_b200``) never does.

Parity status: the reference repository holds no golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the live,
unmodified reference imported from /root/reference in the build container:
``tests/golden/make_golden.py`` produced ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` replays them (bit-level agreement expected,
tolerance 1e-6 relative for safety across torch CPU kernels).
"""
