"""Drop-in for the reference's ``psf_utils.py`` (SyntheticExperiments/psf_utils.py:10-137 and
LRA/psf_utils.py:10-128): ``seed_everything``, ``DatasetCreator``, ``count_params``,
``TrainModel`` (Adding / Temporal Order) and ``TrainPSF`` (LRA).

Same call signatures, same printed quantities, same save rule (state dict written when the test
accuracy beats ``saving_criteria``).  Differences that do not change results: batches are moved
with non-blocking copies and the running loss stays on the device until the end of the epoch
(the reference synchronises with ``.item()`` every step, psf_utils.py:73).
The data-parallel loop lives in ``dp.py``.
"""
from __future__ import annotations

import os
import random
from datetime import datetime

import numpy as np
import torch
from torch.utils.data import Dataset


def seed_everything(seed=1234):
    """Fix every RNG the training scripts touch (psf_utils.py:10-20)."""
    random.seed(seed)
    os.environ['PYTHONHASHSEED'] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed(seed)
    torch.backends.cudnn.deterministic = True


class DatasetCreator(Dataset):
    """(sample, target) pairs over two indexable containers (psf_utils.py:23-41).

    Integer targets are returned as int64: the LRA and Genome copies of this class always cast
    (``LRA/psf_utils.py:37``, ``Genome_Clf/psf_utils.py:37`` - ListOps targets are saved as int32,
    ``listops_preprocessing.py:54``, and CrossEntropyLoss wants int64), the Synthetic copy never does, and its only
    integer targets (Temporal Order) are int64 already; floating targets (Adding, MSE) are left alone.
    ``long_labels=True`` forces the cast, ``False`` forbids it."""

    def __init__(self, data, labels, long_labels=None):
        self.data = data
        self.labels = labels
        self.long_labels = long_labels

    def __getitem__(self, index):
        Y = self.labels[index]
        cast = self.long_labels
        if cast is None:
            cast = torch.is_tensor(Y) and not torch.is_floating_point(Y)
        if cast:
            Y = torch.as_tensor(Y).to(dtype=torch.long)
        return (self.data[index], Y)

    def __len__(self):
        return len(self.labels)


def count_params(net):
    return sum(p.numel() for p in net.parameters() if p.requires_grad)


def _to_device(t):
    return t.cuda(non_blocking=True)


def _train_epoch(net, loader, optimizer, loss, clip_grad_norm=None):
    running = None
    for X, Y in loader:
        X, Y = _to_device(X), _to_device(Y)
        optimizer.zero_grad()
        pred = net(X)
        output = loss(pred.squeeze(), Y)
        output.backward()
        if clip_grad_norm is not None:
            torch.nn.utils.clip_grad_norm_(net.parameters(), max_norm=clip_grad_norm)
        optimizer.step()
        running = output.detach() if running is None else running + output.detach()
    return float(running) / max(len(loader), 1) if running is not None else float('nan')


@torch.no_grad()
def _evaluate(net, loader, loss, problem):
    """Mean loss and accuracy in percent.  Adding counts |pred - Y| < 0.04 (psf_utils.py:102-103)."""
    total, correct, loss_sum = 0, None, None
    for X, Y in loader:
        X, Y = _to_device(X), _to_device(Y)
        pred = net(X)
        l = loss(pred.squeeze(), Y)
        loss_sum = l if loss_sum is None else loss_sum + l
        if problem == 'adding':
            c = (torch.abs(pred.squeeze() - Y) < 0.04).sum()
        else:
            c = pred.max(1)[1].eq(Y).sum()
        correct = c if correct is None else correct + c
        total += Y.size(0)
    if total == 0:
        return float('nan'), float('nan')
    return float(loss_sum) / len(loader), 100.0 * float(correct) / total


def _run(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss, problem, saving_criteria,
         clip_grad_norm=None):
    history = []
    for epoch in range(n_epochs):
        t_start = datetime.now()
        mean_loss = _train_epoch(net, trainloader, optimizer, loss, clip_grad_norm)
        torch.cuda.synchronize()
        seconds = (datetime.now() - t_start).total_seconds()
        print("Epoch {} - Training loss:  {} — Time:  {}sec".format(epoch, mean_loss, seconds))
        record = {"epoch": epoch, "train_loss": mean_loss, "seconds": seconds}
        if epoch % test_freq == 0:
            net.eval()
            val_loss, accuracy_val = _evaluate(net, valloader, loss, problem)
            test_loss, accuracy_test = _evaluate(net, testloader, loss, problem)
            net.train()
            print("Val  loss: {}".format(val_loss))
            print("Test loss: {}".format(test_loss))
            print("Val  accuracy: {}".format(accuracy_val))
            print("Test accuracy: {}".format(accuracy_test))
            print('_' * 40)
            record.update(val_loss=val_loss, test_loss=test_loss, val_acc=accuracy_val, test_acc=accuracy_test)
            if accuracy_test > saving_criteria:
                torch.save(net.state_dict(), '{}_epoch{}_acc{}.pt'.format(problem, epoch, accuracy_test))
        history.append(record)
    return history


def TrainModel(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss, problem,
               saving_criteria, *, clip_grad_norm=None):
    """Synthetic tasks (psf_utils.py:48-137).  Returns the per-epoch records it also prints.
    ``clip_grad_norm`` (keyword only, not in the reference signature): max gradient norm, as the Genome loop applies."""
    return _run(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss, problem,
                saving_criteria, clip_grad_norm)


def TrainPSF(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss, problem,
             saving_criteria, *, clip_grad_norm=None):
    """LRA tasks (LRA/psf_utils.py:48-128): accuracy is always arg-max."""
    return _run(net, trainloader, valloader, testloader, n_epochs, test_freq, optimizer, loss,
                problem if problem != 'adding' else 'order', saving_criteria, clip_grad_norm)
