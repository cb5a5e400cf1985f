"""ORACLE - TEST INFRASTRUCTURE ONLY.

A CPU restatement (plain functional PyTorch, fp32) of the reference algorithm for the hot
path named in BASELINE.json: the fluxion leaf ops, the SD1.5 / SDXL UNet forward, the Euler
denoising step and the SAM ViT image encoder of finegrain-ai/refiners.  Every function cites
the reference file:line it restates (paths relative to /root/reference/src/refiners/).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may import this package - as the checker or the timed CPU baseline, never
as part of the product: nothing under ``refiners_b200/`` imports it.

Parity pin: ``oracle/pin_against_reference.py`` (run where /root/reference is mounted) checks
every function here against the reference's own modules on the same weights and inputs
(fp32, tolerance 1e-5 relative) and writes the committed fixtures under ``tests/golden/``;
``tests/test_oracle_golden.py`` re-checks the oracle against those fixtures anywhere.
"""
