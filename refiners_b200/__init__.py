"""refiners_b200: B200-native execution of refiners' foundation-model forward passes.

Only the hot path named in BASELINE.json lives here: the fluxion Chain/Context/Adapter
mirror (host side, pure Python), the sm_100a kernels behind a C ABI (``csrc/``), and the
models that run on them.  See DESIGN.md.
"""

__version__ = "0.1.0"
