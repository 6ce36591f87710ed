"""Synthesis driver -- eager equivalent of the reference's ``test.py`` (test.py:13-88).

    echo "a prompt to say" | python -m tacotron_b200.test -t nancy

Prompts come from stdin, one per line (test.py:77-78); they are encoded and padded to 140 characters exactly as
``data_input.load_prompts`` does (quirks included), run through ``Tacotron.inference`` in batches of at most 32, every
output is de-normalised with the checkpoint's ``stft_mean`` / ``stft_std`` and inverted with the GPU Griffin-Lim
(test.py:59-64).  Instead of TensorBoard summaries the waveform is written as ``log/<save_path>/test/<n>.wav`` and the
alignment as ``<n>_attention.npy``.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch


def test(model_cls, config, prompts, log=print):
    from . import audio, checkpoint, data_input
    from .train import write_wav
    sr = 24000 if "blizzard" in config.data_path else 16000                              # test.py:15
    meta = data_input.load_meta(config.data_path)                                        # :16
    config.r = meta["r"]          # (the reference reads the module constant audio.r here, :17; meta['r'] is the value it was preprocessed with)
    config.max_decode_iter = audio.maximum_audio_length // (config.r * audio.hop_length)  # tacotron.py:13
    ivocab = meta["vocab"]
    config.vocab_size = len(ivocab)                                                      # :19
    config.num_prompts = len(prompts)                                                    # :23
    model = model_cls(config, None, train=False)                                         # :31
    log("restoring weights")                                                             # :41
    ck = checkpoint.latest_checkpoint(os.path.join("weights", config.save_path))         # :42-45
    if ck is not None:
        stft_mean, stft_std = checkpoint.restore(model, ck)                              # :46-48
    else:
        # no checkpoint of ours: a checkpoint written by the reference itself (TensorFlow V2 format, e.g. the released
        # Nancy weights of download_weights.sh) in the same directory is read directly
        from . import tf_checkpoint
        tf_ck = tf_checkpoint.latest_checkpoint(os.path.dirname(os.path.join("weights", config.save_path)))
        if tf_ck is None:
            raise FileNotFoundError(f"no checkpoint under weights/{config.save_path}")
        extra = checkpoint.import_tf_checkpoint(model, tf_ck)
        stft_mean, stft_std = extra["stft_mean"], extra["stft_std"]
    mean_d = torch.from_numpy(np.asarray(stft_mean, dtype=np.float32)).cuda()
    std_d = torch.from_numpy(np.asarray(stft_std, dtype=np.float32)).cuda()
    out_dir = os.path.join("log", config.save_path, "test")                              # :33
    os.makedirs(out_dir, exist_ok=True)
    n = 0
    waves = []
    for inputs in data_input.load_prompts(prompts, ivocab):                              # :21-22, :50-56
        _, outputs = model.inference(inputs, train=False)
        alignments = model.alignments
        log("saving samples")
        wav = audio.invert_spectrogram(outputs, config.r, stft_mean=mean_d, stft_std=std_d)     # :64, batched on the device
        for b in range(outputs.shape[0]):                                                # :59-70
            text = "".join(ivocab[int(w)] for w in inputs["text"][b].cpu().numpy())
            write_wav(os.path.join(out_dir, f"{n}.wav"), wav[b].cpu().numpy(), sr)
            np.save(os.path.join(out_dir, f"{n}_attention.npy"), alignments[b].cpu().numpy())
            with open(os.path.join(out_dir, f"{n}.txt"), "w") as f:
                f.write(text)
            waves.append(wav[b])
            n += 1
    return waves


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-t", "--train-set", default="nancy")
    args = parser.parse_args(argv)
    prompts = [p for p in sys.stdin.readlines() if len(p) > 0]                           # test.py:77-78
    from .models.tacotron import Config, Tacotron
    config = Config()
    config.data_path = "data/%s/" % args.train_set                                       # :83
    config.save_path = args.train_set + "/tacotron"                                      # :84
    print("Building Tacotron")
    test(Tacotron, config, prompts)


if __name__ == "__main__":
    sys.exit(main())
