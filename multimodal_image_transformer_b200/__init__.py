"""B200-native caption-decoder hot path behind the reference's module interface
(decoder.TransformerDecoder, model.ImageToTextModel, train.train_one_epoch / evaluate,
inference post-processing).  Importing the package does not touch CUDA; constructing a model does,
and fails loudly if libb200decoder.so is missing or the device is not sm_100."""
from . import _lib  # noqa: F401

__all__ = ["config", "utils", "decoder", "model", "train", "inference", "engine", "dp", "ops"]
