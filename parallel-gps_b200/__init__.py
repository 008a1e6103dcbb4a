"""pssgp_b200 — B200-native backend for the temporally-parallel state-space GP inference path of
EEA-sensors/parallel-gps.

The directory is called ``parallel-gps_b200`` (not an importable identifier); load it under the
module name ``pssgp_b200`` with the helper in the repository root::

    from __graft_entry__ import import_package
    pssgp_b200 = import_package()
    from pssgp_b200.kalman.parallel import pkf, pks, pkfs

Layout mirrors the reference package ``pssgp`` for the hot path only:
``kalman.parallel`` (pkf/pks/pkfs), ``kalman.base`` (LGSSM), ``kernels`` (get_sde/get_ssm),
``model`` (StateSpaceGP).  All arithmetic on the path runs in hand-written CUDA kernels for
sm_100a behind the C ABI declared in ``include/pssgp_b200.h``; there is no CPU fallback.
"""
__version__ = "0.1.0"
