"""spef_b200 -- B200-native (sm_100a) implementation of SPEF's batched pose-inference hot path.

Mirrors the reference's Python API for that path (possoj/Spacecraft-Pose-Estimation-Framework):
  modeling.import_model          <- src/modeling/model.py:122-279
  spe.SPEUtils                   <- src/spe/spe_utils.py
  spe.SPEB200 (predict plug-in)  <- src/spe/spe_torch.py (SPETorch)
  tools.evaluation               <- src/tools/evaluation.py
  temporal.Inference/TemporalPDF <- src/temporal/inference.py, src/temporal/pdf_compare.py
  data.FrameTransform            <- Resize + ToTensor of SPEDataset (src/data/utils.py:212-226, src/data/datasets/speed.py:59-62)
Everything numeric runs in libspef_b200.so (include/spef_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
