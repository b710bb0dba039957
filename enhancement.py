#!/usr/bin/env python
"""Command-line enhancement, flag-compatible with the reference's enhancement.py (:28-36):

    python enhancement.py --test_dir noisy/ --enhanced_dir out/ --ckpt model.ckpt --mode score-only \
        [--corrector ald] [--corrector-steps 1] [--snr 0.5] [--N 50]

Extensions: --batch-size (utterances of one padded-frame bucket are enhanced together, ragged lengths inside a batch), --gpus (one process per
GPU via torchrun; files are sharded round-robin, one NCCL weight broadcast, no per-step collective), --seed,
--random-init BACKBONE (run without a checkpoint, for smoke tests), --io-workers (wav reads / writes overlap the GPU work).
wav I/O uses torchaudio like the reference when it has a backend, scipy.io.wavfile otherwise.
"""
import glob
import os
import sys
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
    p.add_argument("--gpus", type=int, default=1, help="GPUs of this node: re-launches itself under torch.distributed.run with one "
                                                        "process per GPU (ignored when already running under torchrun)")
    p.add_argument("--io-workers", type=int, default=4, help="threads reading / writing wav files around the GPU work")
    p.add_argument("--random-init", type=str, default=None, help="backbone name: run with seeded random weights instead of --ckpt")
    return p


def shard(files, rank, world):
    return files[rank::world]


def main(argv=None):
    args = build_parser().parse_args(argv)
    if args.gpus > 1 and "RANK" not in os.environ:      # one process per GPU; 127.0.0.1: the container hostname may not resolve
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + list(sys.argv[1:] if argv is None else argv)
        os.execv(sys.executable, cmd)
    from storm_b200 import DiscriminativeModel, ScoreModel, StochasticRegenerationModel
    from storm_b200.model import checkpoint_kind
    if args.mode == "denoiser-only" and not args.random_init and args.ckpt is None:
        raise SystemExit("--ckpt is required")
    # enhancement.py:47-54: storm -> StochasticRegenerationModel, score-only -> ScoreModel, denoiser-only -> DiscriminativeModel.
    # A StoRM checkpoint (denoiser_net.* / score_net.* keys) is also accepted for denoiser-only: its denoiser runs alone.
    model_cls = {"score-only": ScoreModel, "storm": StochasticRegenerationModel, "denoiser-only": DiscriminativeModel}[args.mode]
    if args.mode == "denoiser-only" and (not args.ckpt or checkpoint_kind(args.ckpt) == "storm"):
        model_cls = StochasticRegenerationModel

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
        for net in ([model.dnn] if hasattr(model, "dnn") else model._nets()):
            parallel.broadcast_weights(net, src=0)

    from storm_b200.parallel import bucket_utterances
    from storm_b200.util import wavio
    files = sorted(glob.glob(os.path.join(args.test_dir, "*.wav")))
    mine = [(i, f) for i, f in enumerate(files)][rank::world]
    model_sr = 16000

    def load(item):
        idx, f = item
        y, sr = wavio.load_wav(f)
        assert sr == model_sr, "You need to make sure sample_sr matches model_sr --> resample to 16kHz"
        return idx, f, y[:1]

    with wavio.AsyncWriter(workers=args.io_workers) as writer:
        # lengths from the wav headers -> batches of one padded-frame bucket (ragged inside the bucket); the next
        # batches are read by worker threads while the GPU enhances the current one, results are written behind it.
        # Every mode runs batched (the reference enhances one file per call, enhancement.py:66-72).
        lengths = [wavio.wav_length(f) for _, f in mine]
        batches = [[mine[i] for i in b] for b in bucket_utterances(lengths, hop_length=model.data_module.hop_length,
                                                                   max_batch=args.batch_size)]
        for group in wavio.Prefetcher(batches, lambda b: [load(it) for it in b], depth=2, workers=args.io_workers):
            ys, ids = [t[2] for t in group], [t[0] for t in group]
            if isinstance(model, StochasticRegenerationModel):
                outs = model.enhance_batch(ys, corrector=args.corrector, N=args.N, corrector_steps=args.corrector_steps, snr=args.snr,
                                           denoiser_only=(args.mode == "denoiser-only"), seed=args.seed, utt_ids=ids,
                                           max_batch=args.batch_size)
            elif isinstance(model, DiscriminativeModel):
                outs = [model.enhance(y) for y in ys] if len({y.numel() for y in ys}) > 1 else list(model.enhance_batch(ys))
            else:
                outs = model.enhance_batch(ys, corrector=args.corrector, N=args.N, corrector_steps=args.corrector_steps, snr=args.snr,
                                           seed=args.seed, utt_ids=ids, max_batch=args.batch_size)
            for (idx, f, _), x_hat in zip(group, outs):
                writer.save(f"{args.enhanced_dir}/{os.path.basename(f)}", x_hat.type(torch.float32).cpu().reshape(1, -1), model_sr)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
