"""Training driver -- eager equivalent of the reference's ``train.py`` (train.py:16-118).

    python -m tacotron_b200.train -t nancy [-r 1] [--steps N]            (one process)
    torchrun --nproc-per-node 8 -m tacotron_b200.train -t nancy         (data parallel, one process per GPU)

Kept from the reference: data_path / save_path conventions (:107-113), meta.pkl supplies r and the vocabulary (:19-22),
lr fed per step starting at ``init_lr`` and multiplied by ``annealing_rate`` every 1000 steps (:60-61, :79-80), the
explosion guard ``loss > 1e8 and global_step > 500`` (:75-77), a checkpoint + an audio sample every SAVE_EVERY steps
(:82-101; the sample is inverted with the GPU Griffin-Lim and written as .wav next to the checkpoint instead of a
TensorBoard audio summary, the alignment as .npy instead of a matplotlib image).  Data parallel ranks draw disjoint
batches of 32 and all-reduce the flat gradient bucket inside ``model.train_step`` (tacotron_b200/optim.py).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

SAVE_EVERY = 5000          # train.py:13
RESTORE_FROM = None        # train.py:14


def write_wav(path, wave, sr):
    from scipy.io import wavfile
    w = np.asarray(wave, dtype=np.float32)
    peak = float(np.max(np.abs(w))) or 1.0
    wavfile.write(path, sr, (w / peak * 0.95 * 32767).astype(np.int16))


def train(model_cls, config, num_steps=1000000, log=print, save_every=SAVE_EVERY):
    from . import audio, checkpoint, data_input
    from .utils import dist as D
    world, rank, local = D.world()
    torch.cuda.set_device(local)
    D.init("nccl")
    sr = 24000 if "vctk" in config.data_path else 16000                                   # train.py:18
    meta = data_input.load_meta(config.data_path)                                        # :19
    config.r = meta["r"]                                                                 # :20
    ivocab = meta["vocab"]
    config.vocab_size = len(ivocab)                                                      # :22
    # every rank must normalise with the SAME statistics (the reference is one process: one 100-utterance sample,
    # data_input.py:54): one seeded draw shared by all ranks instead of the unseeded global numpy RNG per rank
    arrays, names, num_speakers, stft_mean, stft_std = data_input.load_from_npy(
        config.data_path, rng=np.random.RandomState(getattr(config, "data_seed", 0)))                # :26-27
    config.num_speakers = num_speakers                                                   # :29
    batches = data_input.build_dataset(arrays, names, seed=0, shard=(rank, world))       # :35
    model = model_cls(config, None, train=True)                                          # :38
    train_op = model.add_train_op()
    weights_dir = os.path.join("weights", config.save_path)
    if getattr(config, "restore", False):                                                # :49-58
        log("restoring weights")
        ck = (checkpoint.latest_checkpoint(weights_dir) if RESTORE_FROM is None else f"{weights_dir}-{RESTORE_FROM}.npz")
        if ck is not None:
            ck_mean, ck_std = checkpoint.restore(model, ck)
            if ck_mean is not None:                  # de-normalise with the statistics the weights were trained on
                stft_mean, stft_std = ck_mean, ck_std
    mean_d = torch.from_numpy(np.asarray(stft_mean, dtype=np.float32)).cuda()
    std_d = torch.from_numpy(np.asarray(stft_std, dtype=np.float32)).cuda()
    lr = model.config.init_lr                                                            # :60
    annealing_rate = model.config.annealing_rate                                         # :61
    for _ in range(num_steps):                                                           # :63
        inputs = next(batches)
        loss_t = train_op(inputs, lr).reshape(1).clone()                                 # :64-73
        if world > 1:                      # every rank must take the same branch below (a rank that leaves alone hangs the
            torch.distributed.all_reduce(loss_t)   # next gradient all-reduce): the guard and the log use the global loss
        loss = float(loss_t)                                                             # the only host sync of a step
        global_step = model.global_step
        if rank == 0 and global_step % 100 == 0:
            log(f"step {global_step} loss {loss:.1f} lr {lr:.3g} grad-norm {float(model.grad_sumsq.sqrt()):.1f}")
        if loss > 1e8 and global_step > 500:                                             # :75-77 detect gradient explosion
            log("loss exploded")
            break
        if global_step % 1000 == 0:                                                      # :79-80
            lr *= annealing_rate
        if global_step % save_every == 0 and global_step != 0 and rank == 0:             # :82
            log("saving weights")
            fn = checkpoint.save(model, weights_dir, stft_mean, stft_std)                # :85-88
            log("saving sample")                                                         # :90-101
            ideal = audio.invert_spectrogram(inputs["stft"][0], config.r, stft_mean=mean_d, stft_std=std_d)
            sample = audio.invert_spectrogram(model.output[0], config.r, stft_mean=mean_d, stft_std=std_d)
            base = fn[:-4]
            write_wav(base + "_ideal.wav", ideal.cpu().numpy(), sr)
            write_wav(base + "_sample.wav", sample.cpu().numpy(), sr)
            np.save(base + "_attention.npy", model.alignments[0].cpu().numpy())
    return model


def main(argv=None):
    parser = argparse.ArgumentParser()
    parser.add_argument("-t", "--train-set", default="nancy")
    parser.add_argument("-d", "--debug", type=bool, default=False)
    parser.add_argument("-r", "--restore", type=bool, default=False)
    parser.add_argument("--steps", type=int, default=1000000)
    args = parser.parse_args(argv)
    from .models.tacotron import Config, Tacotron
    config = Config()
    config.data_path = "data/%s/" % args.train_set                                       # train.py:108
    config.restore = args.restore
    config.save_path = "debug" if args.debug else "%s/tacotron" % args.train_set         # :110-113
    print("Building Tacotron")
    train(Tacotron, config, num_steps=args.steps)


if __name__ == "__main__":
    sys.exit(main())
