#!/usr/bin/env python
"""Command-line enhancement, flag-compatible with the reference's enhancement.py (:28-36):

    python enhancement.py --test_dir noisy/ --enhanced_dir out/ --ckpt model.ckpt --mode score-only \
        [--corrector ald] [--corrector-steps 1] [--snr 0.5] [--N 50]

Extensions: --batch-size (utterances of equal padded length are enhanced together), --gpus (one process per
GPU via torchrun; files are sharded round-robin, one NCCL weight broadcast, no per-step collective), --seed,
--random-init BACKBONE (run without a checkpoint, for smoke tests).  wav I/O uses torchaudio like the reference.
"""
import glob
import os
from argparse import ArgumentParser

import torch


def build_parser() -> ArgumentParser:
    p = ArgumentParser()
    p.add_argument("--test_dir", type=str, required=True, help="Directory containing your corrupted files to enhance.")
    p.add_argument("--enhanced_dir", type=str, required=True, help="Where to write your cleaned files.")
    p.add_argument("--ckpt", type=str, default=None)
    p.add_argument("--mode", default="score-only", choices=["score-only", "denoiser-only", "storm"])
    p.add_argument("--corrector", type=str, choices=("ald", "langevin", "none"), default="ald", help="Corrector class for the PC sampler.")
    p.add_argument("--corrector-steps", type=int, default=1, help="Number of corrector steps")
    p.add_argument("--snr", type=float, default=0.5, help="SNR value for (annealed) Langevin dynamics.")
    p.add_argument("--N", type=int, default=50, help="Number of reverse steps")
    p.add_argument("--batch-size", type=int, default=8)
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--random-init", type=str, default=None, help="backbone name: run with seeded random weights instead of --ckpt")
    return p


def shard(files, rank, world):
    return files[rank::world]


def main(argv=None):
    args = build_parser().parse_args(argv)
    import torchaudio
    from storm_b200 import ScoreModel, StochasticRegenerationModel
    if args.mode == "denoiser-only" and not args.random_init and args.ckpt is None:
        raise SystemExit("--ckpt is required")
    model_cls = ScoreModel if args.mode == "score-only" else StochasticRegenerationModel   # enhancement.py:47-54

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    os.makedirs(args.enhanced_dir, exist_ok=True)
    if args.ckpt:
        model = model_cls.load_from_checkpoint(args.ckpt, base_dir="", batch_size=1, num_workers=0, kwargs=dict(gpu=False))
    elif args.random_init and args.mode == "score-only":
        model = ScoreModel(backbone=args.random_init, sde="ouve", init_scale=1.0, seed=args.seed)
    elif args.random_init:
        model = StochasticRegenerationModel(backbone_denoiser=args.random_init, backbone_score=args.random_init, sde="ouve",
                                            init_scale=1.0, seed=args.seed)
    else:
        raise SystemExit("--ckpt is required (or --random-init BACKBONE)")
    model.eval(no_ema=False)
    model.cuda()
    if world > 1:   # identical weights everywhere: one broadcast of the flat blob per network
        from storm_b200 import parallel
        for net in ([model.dnn] if args.mode == "score-only" else model._nets()):
            parallel.broadcast_weights(net, src=0)

    files = sorted(glob.glob(os.path.join(args.test_dir, "*.wav")))
    mine = [(i, f) for i, f in enumerate(files)][rank::world]
    model_sr = 16000
    # bucket by padded length so that a batch shares one spectrogram shape
    loaded = []
    for idx, f in mine:
        y, sr = torchaudio.load(f)
        assert sr == model_sr, "You need to make sure sample_sr matches model_sr --> resample to 16kHz"
        loaded.append((idx, f, y[:1]))
    loaded.sort(key=lambda t: t[2].shape[1])
    if args.mode != "score-only":      # StoRM / denoiser-only: one utterance per call like the reference (enhancement.py:66-72)
        for idx, f, y in loaded:
            x_hat = model.enhance(y, corrector=args.corrector, N=args.N, corrector_steps=args.corrector_steps, snr=args.snr,
                                  denoiser_only=(args.mode == "denoiser-only"))
            torchaudio.save(f"{args.enhanced_dir}/{os.path.basename(f)}", x_hat.type(torch.float32).cpu().unsqueeze(0), model_sr)
        loaded = []
    i = 0
    while i < len(loaded):
        L = loaded[i][2].shape[1]
        group = [t for t in loaded[i:i + args.batch_size] if t[2].shape[1] == L]
        ys = torch.cat([t[2] for t in group], dim=0)
        out = model.enhance_batch(ys, corrector=args.corrector, N=args.N, corrector_steps=args.corrector_steps, snr=args.snr,
                                  seed=args.seed, utt_ids=[t[0] for t in group])
        for (idx, f, _), x_hat in zip(group, out):
            torchaudio.save(f"{args.enhanced_dir}/{os.path.basename(f)}", x_hat.type(torch.float32).cpu().unsqueeze(0), model_sr)
        i += len(group)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
