"""ardae -- B200-native AR-DAE hot path behind the reference's model API."""
from .cdae import ConditionalARDAE, MLPGradCARDAE  # noqa: F401
from .ivae import MNISTIPVAE, ToyIPVAE, normal_energy_func  # noqa: F401
from .optim import Adam, RMSprop  # noqa: F401
from .step import TrainStep  # noqa: F401
