"""CPU oracle for the Stable Audio denoising hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the
checker (or as the timed CPU baseline), never as a fallback for the CUDA path.

Contents
--------
* ``dit_oracle``      functional restatement of ``DiffusionTransformer``
                      (reference ``stable_audio_tools/models/dit.py`` and
                      ``models/transformer.py``).
* ``oobleck_oracle``  functional restatement of the Oobleck VAE
                      (``models/autoencoders.py``, ``models/blocks.py``,
                      ``models/bottleneck.py``).
* ``sampler_oracle``  restatement of the k-diffusion 0.1.1 pieces the
                      reference calls (un-vendored third-party dependency;
                      PARITY UNPINNED for those).
* ``ref_shims``       ``sys.modules`` shims that make ``/root/reference``
                      importable in the build container (never on the GPU box).
* ``make_golden``     generates ``tests/golden/*.npz`` from the *real*
                      reference modules (run in the build container only).

Pinning: the reference ships no tests or golden vectors, so the
restatements are pinned against outputs of the reference's own modules run in
the build container (``tests/golden``; regenerate with
``python -m oracle.make_golden``).
"""
