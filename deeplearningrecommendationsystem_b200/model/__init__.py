"""Drop-in replacements for reference model/{lr,mf,ffm,deepfm,afm,nfm,pnn,din,dien,neuralcf,widedeep,deepcross,deepcrossing}.py.

Same constructor signatures, forward signatures, output shapes, state_dict keys/shapes and quirks as the reference
modules, so the reference's scripts/*.py run with only the import line changed.  Embedding lookups, feature
interactions and embedding gradients run in the sm_100a kernels behind include/recsys_b200.h; dense towers stay
on torch.nn.Linear (cuBLAS), as SURVEY.md section 2.2 scopes them.
"""
from .lr import LogisticRegression  # noqa: F401
from .mf import MatrixFactorization  # noqa: F401
from .ffm import FFM  # noqa: F401
from .deepfm import DeepFM  # noqa: F401
from .nfm import NFM  # noqa: F401
from .afm import AFM  # noqa: F401
from .pnn import PNN, DNN, ProductLayers  # noqa: F401
from .din import DIN  # noqa: F401
from .dien import DIEN  # noqa: F401
from .neuralcf import NeuralCF  # noqa: F401
from .widedeep import WideDeep  # noqa: F401
from .deepcross import DeepCross, CrossNetwork, DeepNetwork  # noqa: F401
from .deepcrossing import DeepCrossing, ResidualBlock  # noqa: F401
