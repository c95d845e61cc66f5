"""TEST INFRASTRUCTURE ONLY -- loader for the unmodified Python reference.

This module makes ``/root/reference`` (possoj/Spacecraft-Pose-Estimation-Framework)
importable in the build container so that ``tests/golden/make_goldens.py`` can
run the reference's own code and freeze its outputs as golden fixtures.  It is
never imported by the product package and never runs on the GPU box (the
reference tree does not exist there; ``available()`` returns False).

The reference imports ``brevitas`` unconditionally in every modeling file
(src/modeling/common/pytorch_layers.py:6, src/modeling/head/ursonet.py:6,
src/modeling/backbone/mobilenet_v2.py:6-8, src/modeling/common/quantizers.py:5-10)
although the FP32 path only uses the names inside ``isinstance`` checks
(pytorch_layers.py:18,25).  ``brevitas`` is not installed and there is no
network, so permissive stub modules are registered in ``sys.modules``.  No
reference arithmetic is replaced by the stubs.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("SPEF_REF", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "src", "spe", "spe_utils.py"))


class _Anything:
    """Placeholder class: usable as a base class, in isinstance(), or called."""

    def __init__(self, *a, **k):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (_Anything,), {})
        setattr(self, name, cls)
        return cls


def _install_stubs():
    names = [
        "brevitas", "brevitas.nn", "brevitas.quant", "brevitas.quant.scaled_int",
        "brevitas.quant.shifted_scaled_int", "brevitas.quant.binary", "brevitas.quant.ternary",
        "brevitas.inject", "brevitas.inject.defaults", "brevitas.inject.enum", "brevitas.core",
        "brevitas.core.restrict_val", "brevitas.core.scaling", "brevitas.core.zero_point",
        "brevitas.quant_tensor", "brevitas.export",
    ]
    for n in names:
        if n not in sys.modules:
            m = _StubModule(n)
            m.__path__ = []  # behave like a package
            sys.modules[n] = m
    for n in names:  # `import a.b as x` resolves x via getattr(a, 'b')
        if "." in n:
            parent, child = n.rsplit(".", 1)
            setattr(sys.modules[parent], child, sys.modules[n])
    # quantizers.py:14 evaluates RestrictValueType.LOG_FP at class-definition time
    rv = sys.modules["brevitas.core.restrict_val"]
    rvt = type("RestrictValueType", (), {"LOG_FP": 0, "FP": 1, "INT": 2, "POWER_OF_TWO": 3})
    rv.RestrictValueType = rvt


def load():
    """Put the reference on sys.path (behind stubs) and return its key symbols."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from src.modeling.model import import_model
    from src.spe.spe_utils import SPEUtils
    from src.spe.spe_torch import SPETorch
    from src.spe.classification_utils import OrientationSoftClassification, PositionSoftClassification
    from src.tools.evaluation import evaluation
    from src.tools.utils import RunningAverage
    from src.temporal.pdf_compare import TemporalPDF
    from src.temporal.inference import Inference
    from src.spe.utils import euler2quat, generate_orientation
    ns = types.SimpleNamespace(
        import_model=import_model, SPEUtils=SPEUtils, SPETorch=SPETorch,
        OrientationSoftClassification=OrientationSoftClassification,
        PositionSoftClassification=PositionSoftClassification,
        evaluation=evaluation, RunningAverage=RunningAverage, TemporalPDF=TemporalPDF,
        Inference=Inference, euler2quat=euler2quat, generate_orientation=generate_orientation,
    )
    return ns
