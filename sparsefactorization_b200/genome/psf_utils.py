"""Drop-in for the reference's ``Genome_Clf/psf_utils.py`` (:10-151): ``TrainPSF`` without the ``problem`` argument, gradient
norm clipped to 1.0 on every step (:73), ROC-AUC of the arg-max predictions next to the accuracy (:112,126), int64 targets
(:37).  The reference's save calls have broken format strings (:139-150: four placeholders, two arguments) and would raise;
here the state dict is written as ``genome_epoch{e}_acc{a}.pt``."""
from __future__ import annotations

from datetime import datetime

import torch

from .. import psf_utils as _base
from ..psf_utils import count_params, seed_everything  # noqa: F401

__all__ = ["seed_everything", "DatasetCreator", "count_params", "TrainPSF", "roc_auc"]


class DatasetCreator(_base.DatasetCreator):
    def __init__(self, data, labels):
        super().__init__(data, labels, long_labels=True)


def roc_auc(targets: torch.Tensor, scores: torch.Tensor) -> float:
    """Area under the ROC curve of binary ``targets`` for ``scores`` (ties get the mid-rank: sklearn's roc_auc_score)."""
    t = targets.reshape(-1).double().cpu()
    s = scores.reshape(-1).double().cpu()
    n_pos, n_neg = float((t == 1).sum()), float((t != 1).sum())
    if n_pos == 0 or n_neg == 0:
        return float("nan")
    order = torch.argsort(s)
    ranks = torch.empty_like(s)
    ranks[order] = torch.arange(1, s.numel() + 1, dtype=torch.float64)
    uniq, inv = torch.unique(s, return_inverse=True)          # mid-ranks for tied scores
    sums = torch.zeros(uniq.numel(), dtype=torch.float64).index_add_(0, inv, ranks)
    cnts = torch.zeros(uniq.numel(), dtype=torch.float64).index_add_(0, inv, torch.ones_like(ranks))
    ranks = (sums / cnts)[inv]
    return float((ranks[t == 1].sum() - n_pos * (n_pos + 1) / 2) / (n_pos * n_neg))


@torch.no_grad()
def _evaluate(net, loader, loss):
    total, correct, loss_sum, tg, pr = 0, None, None, [], []
    for X, Y in loader:
        X, Y = X.cuda(non_blocking=True), Y.cuda(non_blocking=True)
        pred = net(X)
        l = loss(pred.squeeze(), Y)
        loss_sum = l if loss_sum is None else loss_sum + l
        predicted = pred.max(1)[1]
        c = predicted.eq(Y).sum()
        correct = c if correct is None else correct + c
        total += Y.size(0)
        tg.append(Y)
        pr.append(predicted)
    if total == 0:
        return float("nan"), float("nan"), float("nan")
    return float(loss_sum) / len(loader), 100.0 * float(correct) / total, roc_auc(torch.cat(tg), torch.cat(pr))


def TrainPSF(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss, saving_criteria):
    history = []
    for epoch in range(n_epochs):
        t_start = datetime.now()
        print(len(trainloader))
        mean_loss = _base._train_epoch(net, trainloader, optimizer, loss, clip_grad_norm=1.0)
        torch.cuda.synchronize()
        seconds = (datetime.now() - t_start).total_seconds()
        print("Epoch {} - Training loss:  {} \u2014 Time:  {}sec".format(epoch, mean_loss, seconds))
        record = {"epoch": epoch, "train_loss": mean_loss, "seconds": seconds}
        if epoch % test_freq == 0:
            net.eval()
            val_loss, acc_val, auc_val = _evaluate(net, valloader, loss)
            test_loss, acc_test, auc_test = _evaluate(net, testloader, loss)
            net.train()
            print("Val  loss: {}".format(val_loss))
            print("Test loss: {}".format(test_loss))
            print("Val  accuracy: {}".format(acc_val))
            print("Test accuracy: {}".format(acc_test))
            print("Val  ROCAUC: {}".format(100. * auc_val))
            print("Test ROCAUC: {}".format(100. * auc_test))
            print('_' * 40)
            record.update(val_loss=val_loss, test_loss=test_loss, val_acc=acc_val, test_acc=acc_test,
                          val_rocauc=auc_val, test_rocauc=auc_test)
            if acc_test > saving_criteria:
                torch.save(net.state_dict(), 'genome_epoch{}_acc{}.pt'.format(epoch, acc_test))
        history.append(record)
    return history
