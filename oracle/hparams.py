"""Plain-Python mirror of the hot-path hyper-parameters (TEST INFRASTRUCTURE).

Follows /root/reference/configs/hparams.py:233-348 (LJHPS) and :351-474 (DataBakerHPS);
only the values consumed by models/models.py:16-65 are kept, without the ``tf.`` handles.
"""


class _Common:
    latent_dim = 128
    output_dim = 80
    final_reduction_factor = 2
    max_reduction_factor = 5
    mel_text_len_ratio = 5.59


class _Encoder:
    vocab_size = 43
    embd_dim = 512
    n_conv = 3
    pre_hidden = 512
    conv_kernel = 5
    pre_drop_rate = 0.1
    pos_drop_rate = 0.1
    bn_before_act = False
    n_blk = 4
    attention_dim = 256
    attention_heads = 4
    attention_temperature = 1.0
    ffn_hidden = 1024


class _Decoder:
    nblk = 2
    attention_dim = 256
    attention_heads = 4
    ffn_hidden = 1024
    attention_temperature = 1.0
    post_n_conv = 5
    post_conv_filters = 256
    post_conv_kernel = 5
    post_drop_rate = 0.2


class _Posterior:
    pre_hidden = 256
    pos_drop_rate = 0.2
    pre_drop_rate = 0.5
    nblk = 2
    attention_dim = 256
    attention_heads = 4
    temperature = 1.0
    ffn_hidden = 1024


class _Prior:
    n_blk = 6
    n_transformer_blk = 2
    attention_dim = 256
    attention_heads = 4
    temperature = 1.0
    ffn_hidden = 1024
    inverse = False


class _Train:
    random_seed = 123456
    train_batch_size = 32
    num_samples = 1
    length_weight = 1.0
    kl_weight_init = 1e-5
    kl_weight_end = 1e-5
    learning_rate = 1.25e-4
    reduction_factors = [5, 4, 3, 2]
    reduce_interval = [0, 200, 400, 600]


class LJHPS:
    """configs/hparams.py:233-348"""
    name = "ljspeech"
    Train = _Train
    Common = _Common
    Encoder = _Encoder
    Decoder = _Decoder
    Posterior = _Posterior
    Prior = _Prior
    num_mels = 80


class _DBCommon(_Common):
    mel_text_len_ratio = 4.21


class _DBEncoder(_Encoder):
    vocab_size = 39


class _DBTrain(_Train):
    random_seed = 12


class DataBakerHPS(LJHPS):
    """configs/hparams.py:351-474 (differs in vocab, ratio, seed, audio only)"""
    name = "databaker"
    Train = _DBTrain
    Common = _DBCommon
    Encoder = _DBEncoder


class TinyHPS(LJHPS):
    """Shrunk architecture for fast unit tests / small golden files (NOT a reference config)."""
    name = "tiny"

    class Common(_Common):
        latent_dim = 32

    class Encoder(_Encoder):
        embd_dim = 64
        pre_hidden = 64
        n_blk = 1
        n_conv = 2
        attention_dim = 64
        attention_heads = 2
        ffn_hidden = 128

    class Decoder(_Decoder):
        nblk = 1
        attention_dim = 64
        attention_heads = 2
        ffn_hidden = 128
        post_n_conv = 2
        post_conv_filters = 64

    class Posterior(_Posterior):
        pre_hidden = 64
        nblk = 1
        attention_dim = 64
        attention_heads = 2
        ffn_hidden = 128

    class Prior(_Prior):
        n_blk = 2
        n_transformer_blk = 1
        attention_dim = 64
        attention_heads = 2
        ffn_hidden = 128
