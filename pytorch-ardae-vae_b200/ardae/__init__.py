"""ardae -- B200-native AR-DAE hot path behind the reference's model API."""
from .cdae import ConditionalARDAE, MLPGradCARDAE  # noqa: F401
