"""hcmvs_b200 — B200-native PatchMatch depth estimation, depth-map filtering and fusion
(the HC-MVS / OpenMVS ``libs/MVS`` dense-reconstruction hot path) behind a C ABI.

Python here is plumbing only: ctypes bindings over ``libhcmvs_b200.so`` (CUDA, sm_100a) and
``libhcmvs_host.so`` (scene synthesis / file formats). There is no CPU fallback: importing
``hcmvs_b200.api`` raises if the CUDA library is missing.
"""
__version__ = "0.1.0"
