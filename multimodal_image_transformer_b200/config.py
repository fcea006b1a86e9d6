"""Hyper-parameters of the caption decoder hot path, under the names the reference's config.py
uses (reference config.py:10-145) so that callers written against it keep working.  Only the
constants the hot path reads are kept; dataset / wandb / hub settings are out of scope."""
import torch

DEVICE = "cuda" if torch.cuda.is_available() else "cpu"
RANDOM_SEED = 42

ENCODER_MODEL_NAME = "google/vit-base-patch16-224-in21k"
IMAGE_PROCESSOR_NAME = "google/vit-base-patch16-224-in21k"

VOCAB_SIZE = 10000
MAX_SEQ_LEN = 100
DECODER_EMBED_DIM = 512
DECODER_LAYERS = 6
DECODER_HEADS = 8
DECODER_FF_DIM = 2048
DECODER_DROPOUT = 0.1
PROJECTION_DIM = 512

BATCH_SIZE = 32
NUM_EPOCHS = 20
LEARNING_RATE = 1e-4
WEIGHT_DECAY = 1e-5
GRAD_CLIP_VALUE = 5.0
ADAM_BETA1 = 0.9
ADAM_BETA2 = 0.98
ADAM_EPS = 1e-9
WARMUP_STEPS = 0
LOG_INTERVAL = 50

PAD_TOKEN, START_TOKEN, END_TOKEN, UNK_TOKEN = "<PAD>", "<START>", "<END>", "<UNK>"
PAD_TOKEN_ID = 0
START_TOKEN_ID = 1
END_TOKEN_ID = 2
UNK_TOKEN_ID = 3
BEAM_SIZE = 3

# B200-specific additions (not in the reference)
MEMORY_MODE = "cls"        # "cls": one memory token per image (reference model.py:141,151);
                           # "patch": all encoder tokens (50 / 197 / 257) feed cross-attention
