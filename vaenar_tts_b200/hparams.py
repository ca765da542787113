"""Hyper-parameters of the mel-synthesis path with the reference's own nesting
(``hps.Encoder.Transformer.embd_dim`` ...), mirroring /root/reference/configs/hparams.py:233-348 (LJHPS)
and :351-474 (DataBakerHPS) without the ``tf.`` handles.  ``VAENAR(hps)`` also accepts the reference's
original hparams classes: only plain numeric attributes are read."""


class LJHPS:
    class Train:
        random_seed = 123456
        epochs = 2000
        train_batch_size = 32
        test_batch_size = 8
        num_samples = 1
        length_weight = 1.
        kl_weight = 1.
        kl_weight_init = 1e-5
        kl_weight_increase_epoch = 1
        kl_weight_end = 1e-5
        learning_rate = 1.25e-4
        reduction_factors = [5, 4, 3, 2]
        reduce_interval = [0, 200, 400, 600]

    class Audio:                      # configs/hparams.py:266-282; consumed by vaenar_tts_b200.audio (mel inversion)
        num_mels = 80
        num_freq = 1025
        min_mel_freq = 0.
        max_mel_freq = 8000.
        sample_rate = 22050
        frame_length_sample = 1024
        frame_shift_sample = 256
        preemphasize = 0.97
        min_level_db = -100.0
        ref_level_db = 20.0
        max_abs_value = 1
        symmetric_specs = False
        griffin_lim_iters = 60
        power = 1.5
        center = True

    class Common:
        latent_dim = 128
        output_dim = 80
        final_reduction_factor = 2
        max_reduction_factor = 5
        mel_text_len_ratio = 5.59

    class Encoder:
        class Transformer:
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

    class Decoder:
        class Transformer:
            nblk = 2
            attention_dim = 256
            attention_heads = 4
            ffn_hidden = 1024
            attention_temperature = 1.
            post_n_conv = 5
            post_conv_filters = 256
            post_conv_kernel = 5
            post_drop_rate = 0.2

    class Posterior:
        class Transformer:
            pre_hidden = 256
            pos_drop_rate = 0.2
            pre_drop_rate = 0.5
            nblk = 2
            attention_dim = 256
            attention_heads = 4
            temperature = 1.0
            ffn_hidden = 1024

    class Prior:
        class Transformer:
            n_blk = 6
            n_transformer_blk = 2
            attention_dim = 256
            attention_heads = 4
            temperature = 1.0
            ffn_hidden = 1024
            inverse = False


class DataBakerHPS(LJHPS):
    class Train(LJHPS.Train):
        random_seed = 12

    class Audio(LJHPS.Audio):         # configs/hparams.py:384-400
        sample_rate = 16000
        frame_length_sample = 800
        frame_shift_sample = 200
        min_level_db = -115.

    class Common(LJHPS.Common):
        mel_text_len_ratio = 4.21

    class Encoder:
        class Transformer(LJHPS.Encoder.Transformer):
            vocab_size = 39
