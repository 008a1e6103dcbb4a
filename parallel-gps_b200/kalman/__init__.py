from .base import LGSSM  # noqa: F401
from .parallel import pkf, pks, pkfs  # noqa: F401
from .sequential import kf, ks, kfs  # noqa: F401
