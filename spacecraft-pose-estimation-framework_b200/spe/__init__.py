from .classification_utils import OrientationSoftClassification, PositionSoftClassification  # noqa: F401
from .spe_utils import SPEUtils  # noqa: F401
from .spe_b200 import SPEB200  # noqa: F401
