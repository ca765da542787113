"""The caller of the path in the reference's ``inference.py``: the length-predictor-driven
``test_step`` loop with real-time-factor accounting (inference.py:148-168) and the per-utterance mel writer
(audio/utils.py:16-22).  Host plumbing around ``VAENAR.test_step``; all arithmetic stays in the CUDA path."""
import os
import time

import numpy as np
import torch


def write_mels(save_dir, step, mel_batch, mel_lengths, ids, prefix=""):
    """``TestUtils.write_mels`` (audio/utils.py:16-22): one ``{prefix}-{id}-{step}.npy`` per utterance, cropped to its
    length.  Returns the file names."""
    os.makedirs(save_dir, exist_ok=True)
    mel_batch = np.asarray(mel_batch.detach().cpu() if hasattr(mel_batch, "detach") else mel_batch)
    mel_lengths = np.asarray(mel_lengths.detach().cpu() if hasattr(mel_lengths, "detach") else mel_lengths)
    names = []
    for i in range(mel_batch.shape[0]):
        mel = mel_batch[i][:int(mel_lengths[i]), :]
        idx = ids[i].decode("utf-8") if isinstance(ids[i], bytes) else ids[i]
        name = os.path.join(save_dir, "{}-{}-{}.npy".format(prefix, idx, step))
        np.save(name, mel)
        names.append(name)
    return names


def inference_test(model, batches, frame_shift_sample, sample_rate, temperature=0.0, save_dir=None, ckpt_step=0,
                   write_mel_files=False, warmup=True):
    """The synthesis loop of inference.py:145-168.  ``batches`` yields ``(fids, texts, t_lengths)`` (or the 5-tuples of the
    input pipeline: fids, texts, mels, t_lengths, m_lengths).  Each ``test_step`` is timed with the wall clock INCLUDING
    the device synchronisation (the reference's ``.numpy()`` calls imply it); durations = predicted frames * frame shift /
    sample rate.  Returns dict(time_consumed, durations, average_rtf, n_utterances)."""
    batches = list(batches)

    def unpack(b):
        return (b[0], b[1], b[3]) if len(b) == 5 else b
    if warmup and batches:                       # "tf.function initialization" pass of inference.py:146-147
        _, texts, t_l = unpack(batches[0])
        model.test_step(texts, t_l, temperature=temperature)
        torch.cuda.synchronize()
    time_consumed, durations, n = 0.0, 0.0, 0
    for b in batches:
        fids, texts, t_l = unpack(b)
        t0 = time.time()
        mel, pred_m_lens, _ = model.test_step(texts, t_l, temperature=temperature, return_alignments=False)
        lens = pred_m_lens.cpu().numpy()         # device -> host: synchronises, like the reference's .numpy()
        time_consumed += time.time() - t0
        durations += float(np.sum(lens)) * frame_shift_sample / sample_rate
        n += len(lens)
        if write_mel_files and save_dir is not None:
            write_mels(save_dir, ckpt_step, mel, np.minimum(lens, mel.shape[1]), fids, prefix="prior")
    return {"time_consumed": time_consumed, "durations": durations,
            "average_rtf": time_consumed / durations if durations > 0 else float("nan"), "n_utterances": n}
