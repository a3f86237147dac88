"""B200-native (sm_100a) Conformer encoder: drop-in for the encoder of
Lingeng56/conformer-pytorch-lightning (src/encoder.py, encoder_layer.py, attention.py,
convolution.py, feedforward.py, utils.py mask helpers), executed by hand-written CUDA kernels
behind the C-ABI in include/cfm_b200.h.  No CPU fallback: CPU tensors raise."""
from .attention import (MultiHeadSelfAttentionModule, PositionalEncoding, RelativeMultiHeadSelfAttentionModule,
                        RelativePositionalEncoding)
from .convolution import ConvolutionModule, ConvolutionSubSampling
from .encoder import ConformerEncoder
from .encoder_layer import ConformerEncoderLayer
from .feedforward import PositionwiseFeedForwardModule
from .pipeline import EncoderPipeline
from .ctc import CTCDecoder, CTCGreedyHead
from .frontend import Fbank, GlobalCMVN, load_cmvn
from .transducer import RNNPredictor, TransducerJoint, basic_greedy_search, rnnt_loss
from . import ddp
from . import optim
from .optim import FlatAdam
from .utils import make_attn_mask, make_pad_mask, subsequent_chunk_mask

__all__ = ["FlatAdam", "optim", "ConformerEncoder", "ConformerEncoderLayer", "RelativeMultiHeadSelfAttentionModule",
           "MultiHeadSelfAttentionModule", "RelativePositionalEncoding", "PositionalEncoding", "ConvolutionModule",
           "ConvolutionSubSampling", "PositionwiseFeedForwardModule", "make_pad_mask", "make_attn_mask",
           "subsequent_chunk_mask", "EncoderPipeline", "CTCGreedyHead", "CTCDecoder", "ddp", "Fbank", "GlobalCMVN", "load_cmvn", "TransducerJoint", "RNNPredictor",
           "rnnt_loss", "basic_greedy_search"]
