"""End-to-end accuracy parity (north-star: "end-to-end test accuracy must fall within the reference's run-to-run
spread"): the full CLI run — ``train.main`` with the reference's flags: training steps, bandit updates, StepLR per epoch,
per-epoch validation with the stochastic sampler, best-val checkpoint reload, layer-wise full-neighbour inference,
micro-F1 (``train_lightning.py:562-705``) — on a planted-partition synthetic graph of a dataset shape, k = 5 runs,
against k = 5 runs of the same procedure restated on the CPU oracle (``tests/util.oracle_fit``).  The synthetic
graphs of BASELINE.json carry random labels (nothing to learn); ``synthetic:<name>:planted`` plants communities that
drive edges, features and labels, so that sampling quality shows up in the accuracy.

The runs differ by their seeds (model init, batch order, Philox stream, dropout masks — the device draws dropout
from Philox, the oracle from torch's generator, so the two sides are independent samples of the same procedure).
Asserted: the device's mean test micro-F1 lies inside the oracle's [min, max] over its runs (widened by one standard
error of the device mean), and both are far above chance."""
import os

import pytest
import torch

from tests.util import oracle_fit, record

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,batch,fan,n_steps", [("cora", 32, "512,256,128", 60), ("pubmed", 32, "512,256,128", 12)])
def test_cli_run_accuracy_within_oracle_spread(native_lib, tmp_path, shape, batch, fan, n_steps):
    from bliss_gnn_b200 import train
    from bliss_gnn_b200.graph import normalized_edata, synthetic_graph
    k = 5
    assert torch.cuda.is_available()
    argv = ["--dataset", f"synthetic:{shape}:planted", "--model", "sage", "--sampler", "poisson-bandit", "--batch-size",
            str(batch), "--fan-out", fan, "--num-steps", str(n_steps), "--k-runs", str(k), "--logdir", str(tmp_path),
            "--precision", "highest", "--gpu", "0"]
    dev_runs = [r["Test"] for r in train.main(argv)]
    # the best-val checkpoint and the metrics log of the LAST run exist (ModelCheckpoint / logger, :621-647)
    versions = sorted(os.listdir(os.path.join(tmp_path, os.listdir(tmp_path)[0])))
    assert len(versions) == k
    last = os.path.join(tmp_path, os.listdir(tmp_path)[0], versions[-1])
    assert os.listdir(os.path.join(last, "checkpoints")) and os.path.exists(os.path.join(last, "metrics.jsonl"))
    g = synthetic_graph(shape, seed=0, device=torch.device("cuda:0"), planted=True).to("cpu")
    g.edata["w"] = normalized_edata(g)
    ora_runs = [oracle_fit(g, [int(f) for f in fan.split(",")], batch, 256, n_steps, seed=run) for run in range(k)]
    dm, om = sum(dev_runs) / k, sum(ora_runs) / k
    se = (sum((x - dm) ** 2 for x in dev_runs) / (k * (k - 1))) ** 0.5
    record(f"accuracy-{shape}", {"device_runs": dev_runs, "oracle_runs": ora_runs, "device_mean": dm, "oracle_mean": om})
    chance = 1.0 / g.n_classes
    assert dm > chance + 0.1 and om > chance + 0.1, (dev_runs, ora_runs)
    assert min(ora_runs) - se <= dm <= max(ora_runs) + se, (dev_runs, ora_runs)
