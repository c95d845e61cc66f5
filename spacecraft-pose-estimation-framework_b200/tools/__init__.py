from .evaluation import evaluation, mad  # noqa: F401
from .utils import RunningAverage, AverageMeter  # noqa: F401
