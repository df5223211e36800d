#!/usr/bin/env python
"""Drop-in script name of the reference (``train_lightning.py``): the same 30 flags
(``--sampler poisson-bandit|bandit|ladies|poisson-ladies --model sage|gcn|gat --fan-out … --eta …``), plus
``--seed`` / ``--normalize``; ``--dataset`` takes ``toy``, a dataset name or ``synthetic:<name>[:scale]``.
Data-parallel: ``torchrun --nproc-per-node N train_lightning.py …`` (one process per GPU)."""
from bliss_gnn_b200.train import main

if __name__ == "__main__":
    main()
