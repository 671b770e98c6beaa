"""Mirror of the one function of /root/reference/SimCLR/Model_Util.py that consumes the hot path's
outputs: ``top_k_accuracy`` (Model_Util.py:104-113), used by validate() for the contrastive top-1 /
top-5 accuracy (Contrastive_Learning.py:867-868).  Everything else in the reference's Model_Util
(LR schedule, optimiser factory, checkpointing) is training plumbing and out of scope."""
import torch


def top_k_accuracy(preds, target, k):
    """Same name, argument order and return value (0-dim float tensor) as the reference.

    Two input forms:
      * the reference's: ``preds`` (bsz, C) scores, ``target`` (bsz,) class indices or (bsz, C)
        one-hot (Model_Util.py:106-109);
      * the fused one: ``preds`` = int32 ``pos_rank`` (bsz,) from
        ``contrastive_loss(..., fused_topk=True)`` and ``target`` = None.  The positive is inside the
        top k exactly when fewer than k keys beat it, so the accuracy is mean(pos_rank < k).
    """
    if target is None:
        if preds.dim() != 1 or preds.dtype not in (torch.int32, torch.int64):
            raise TypeError("with target=None, preds must be the int pos_rank vector of "
                            "contrastive_loss(..., fused_topk=True)")
        return (preds < int(k)).sum() / (preds.shape[0] + 0.0)
    a = torch.transpose(torch.topk(preds, k=k, dim=1)[1], 0, 1)
    b = target if target.dim() == 1 else torch.argmax(target, dim=1)
    d = torch.any(a == b, dim=0)
    return torch.sum(d) / (d.shape[0] + 0.0)
