"""CLI -- drop-in for ``python -m clane`` (/root/reference/clane/__main__.py:15-126).

Same flags (``--data_root --output_root --config_file --save_history --num_workers --gpu``),
same YAML sections (``graph``, ``similarity{method,kwargs}``, ``embedder``) splatted as kwargs,
same outputs (``output_root/{outer}/Z_{sweep}.npy`` with ``--save_history``, ``output_root/Z.npy``).
Also accepts the README form ``clane embedding --...`` (the leading word is optional).
The update always runs on the GPU; ``--gpu`` is accepted for compatibility.
"""
from __future__ import annotations

import argparse
import sys
from pathlib import Path

import numpy as np
import torch
import yaml

from . import similarity
from .embedder import Embedder, IterativeEmbedder
from .graph import Graph


def embedding(args):
    print('[Embedding]', end='\n')

    config_path = args.config_file.absolute()
    if not config_path.exists():
        raise FileNotFoundError(f"Config file not found. {config_path}")
    with open(config_path, 'r') as config_io:
        hparams = yaml.load(config_io, Loader=yaml.FullLoader)

    device = torch.device('cuda')
    g = Graph(data_root=args.data_root, **hparams["graph"])

    print("Graph Loaded.")
    print(f" - {len(g)} vertices")
    print(f" - {len(g.E)} edges")
    print(" - Content Embeddings:")
    print(f"     - dim : {g.d:3d}")
    print(f"     - mean: {g.X.mean():5.2f}")
    print(f"     - std : {g.X.std():5.2f}")

    method = hparams["similarity"]["method"]
    try:
        similarity_cls = getattr(similarity, method)
    except AttributeError:
        raise AttributeError(f'Given similarity method {method} not found.')
    similarity_measure = similarity_cls(**hparams['similarity']['kwargs'])

    common = dict(graph=g, similarity_measure=similarity_measure, device=device, save_history=args.save_history)
    if hasattr(similarity_measure, 'parameters'):
        embedder = IterativeEmbedder(num_workers=args.num_workers, **common, **hparams["embedder"])
    else:
        embedder = Embedder(**common, **hparams["embedder"])
    args.output_root.mkdir(parents=True, exist_ok=True)
    if args.save_history:
        # output_root/{outer}/Z_{sweep}.npy (__main__.py:73-82 of the reference) is written while the sweeps run:
        # at products shape one sweep's Z is ~1 GB, and the reference keeps every one of them in memory until the end
        embedder.history_root = args.output_root
    embedder.iterate()

    print("Saving the results.")
    np.save(args.output_root.joinpath('Z.npy'), g.Z.cpu().numpy())
    print(f"The embeddings are stored in {args.output_root.joinpath('Z.npy').absolute()}.")


def get_parser():
    parser = argparse.ArgumentParser(prog="clane")
    parser.add_argument("--data_root", type=Path, help="Path to the data root directory.")
    parser.add_argument("--output_root", type=Path, help="Path to the root for the experiment results to be stored.")
    parser.add_argument("--config_file", type=Path, help="Path to the training configuration yaml file.")
    parser.add_argument("--save_history", action='store_true',
                        help="If true, it saves the embeddings for every iteration.")
    parser.add_argument("--num_workers", type=int, default=0)
    parser.add_argument("--gpu", action='store_true')
    return parser


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if argv and argv[0] == "embedding":   # README form: `clane embedding --data_root ...`
        argv = argv[1:]
    args = get_parser().parse_args(argv)
    embedding(args)


if __name__ == "__main__":
    main()
