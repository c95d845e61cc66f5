from .pdf_compare import TemporalPDF  # noqa: F401
from .inference import Inference  # noqa: F401
