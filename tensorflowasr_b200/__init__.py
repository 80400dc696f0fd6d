"""tensorflowasr_b200 — B200-native (sm_100a) implementation of TensorFlowASR's data-parallel training hot path.

Public surface mirrors the reference (TensorSpeech/TensorFlowASR): `FeatureExtraction` / `SpeechFeaturizer`,
`RnntLoss`, `CtcLoss`, `rnnt_loss`, the `schemas` NamedTuples, and the model wrappers, with torch CUDA tensors in
place of tf.Tensors.  Compute runs in hand-written CUDA behind the C ABI of include/tfasr_b200.h.
"""
from . import schemas  # noqa: F401
from .features import FeatureExtraction, SpeechFeaturizer  # noqa: F401
from .losses import BaseLoss, CtcLoss, RnntLoss, ctc_loss, rnnt_loss  # noqa: F401

__version__ = "0.1.0"
from . import ops  # noqa: F401,E402
from .base_model import BaseModel, TransformerSchedule, make_optimizer  # noqa: F401,E402
from .encoder import ConformerEncoder, Conv2dSubsampling  # noqa: F401,E402
from .transducer import Conformer, Transducer, TransducerJoint, TransducerPrediction, conformer_small_config  # noqa: F401,E402
from .ctc_models import (ContextNet, ContextNetEncoder, CtcModel, DeepSpeech2, DeepSpeech2Encoder, contextnet_config,  # noqa: F401,E402
                         deepspeech2_base_config)
