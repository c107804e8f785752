"""Ensemble statistics used by ``enhance(ensemble=..., ensemble_stat="signal_median")``
(reference ``utils/stats.py:22-66``).  Host-side torch glue on whatever device the samples are
on: it runs once per call on E x B finished clips and is not part of the kernel path."""
import torch


def signal_median(signal):
    """signal: (ensemble, batch, ...) -> (batch, ...): ONE ensemble member per batch row.

    Upstream's rule, kept as is (including its use of the winning *rank* as a member index):
    sort the members sample by sample, note at which rank the member whose index is closest to
    n/2 sits (first rank on ties), and return, per row, the member whose index equals the rank
    that occurred most often."""
    shape = signal.shape
    flat = signal.flatten(start_dim=2)                       # (E, B, S)
    n = flat.shape[0]
    _, order = flat.sort(dim=0)
    _, rank = (order - n / 2).abs().min(dim=0)               # (B, S)
    counts = torch.nn.functional.one_hot(rank, n).sum(dim=1)  # (B, E)
    select = counts.argmax(dim=1)                            # (B,)
    rows = torch.arange(flat.shape[1], device=flat.device)
    return flat[select, rows].reshape(shape[1:])
