"""spectre_b200: B200-native DG evolution right-hand side (ScalarWave and
GeneralizedHarmonic) behind SpECTRE's operator surface.  See DESIGN.md."""
from . import lib  # noqa: F401
