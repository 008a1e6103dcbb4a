from .base import LGSSM  # noqa: F401
from .parallel import pkf, pks, pkfs  # noqa: F401
