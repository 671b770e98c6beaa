"""Mirror of the functions of /root/reference/SimCLR/Model_Util.py that sit on either side of the hot
path in the training / validation loops:

* ``top_k_accuracy`` (Model_Util.py:104-113) -- consumer of the loss's outputs in validate()
  (Contrastive_Learning.py:867-868);
* ``learning_rate_schedule`` (Model_Util.py:9-39, helpers :50-61) -- called between the loss and
  ``loss.backward()`` in train() (Contrastive_Learning.py:693-694);
* ``get_optimizer`` (Model_Util.py:68-88) incl. the 'lars' choice, whose ``apex.parallel.LARC`` wrapper is
  not installable here and is restated as plain PyTorch (``LARC`` below).

Same names, argument meaning and error behaviour as the reference, so the harness of configs[3]
(tools/simclr_step.py, tools/convergence_parity.py) reads like the reference's loop.  Checkpointing and
plotting (Model_Util.py:95-100, 132-148) are out of scope."""
import math

import torch
import torch.optim as optim


def top_k_accuracy(preds, target, k):
    """Same name, argument order and return value (0-dim float tensor) as the reference.

    Two input forms:
      * the reference's: ``preds`` (bsz, C) scores, ``target`` (bsz,) class indices or (bsz, C)
        one-hot (Model_Util.py:106-109);
      * the fused one: ``preds`` = int32 ``pos_rank`` (bsz,) from
        ``contrastive_loss(..., fused_topk=True)`` and ``target`` = None.

    Both reduce to one definition: the target is inside the top k exactly when fewer than k scores of
    its row beat it, so accuracy = mean(rank_of_target < k).  (Exact ties count for the target;
    ``torch.topk`` in the reference breaks them arbitrarily.)
    """
    if target is None:
        if preds.dim() != 1 or preds.dtype not in (torch.int32, torch.int64):
            raise TypeError("with target=None, preds must be the int pos_rank vector of "
                            "contrastive_loss(..., fused_topk=True)")
        rank_of_target = preds
    else:
        cls = target if target.dim() == 1 else target.argmax(dim=1)
        rank_of_target = (preds > preds.gather(1, cls.view(-1, 1))).sum(dim=1)
    return (rank_of_target < int(k)).sum() / float(rank_of_target.shape[0])


def _cosine_decay(learning_rate, global_step, decay_steps, alpha=0.0):
    """Half-cosine from ``learning_rate`` to ``alpha * learning_rate`` over ``decay_steps`` (Model_Util.py:50-54)."""
    t = min(global_step, decay_steps) / decay_steps
    return learning_rate * ((1.0 - alpha) * 0.5 * (1.0 + math.cos(math.pi * t)) + alpha)


def _get_train_steps(num_examples, train_epochs, train_batch_size):
    """Model_Util.py:58-60."""
    return num_examples * train_epochs // train_batch_size + 1


def learning_rate_schedule(arguments):
    """Linear warm-up then cosine decay, written into every param group (Model_Util.py:9-39).

    ``arguments`` is the reference's dict: 'optimizer', 'warmup_epochs', 'num_examples', 'batch_size',
    'world_size', 'learning_rate_scaling' ('linear': base * global_batch / 256, 'sqrt': base *
    sqrt(global_batch); anything else raises ValueError), 'base_learning_rate', 'train_epochs'.  The step
    count is read from the optimiser state of the last parameter of group 0 (1 before the first step),
    exactly like the reference, so Adam-family optimisers drive it and plain SGD stays at step 1."""
    opt = arguments['optimizer']
    state = opt.state[opt.param_groups[0]["params"][-1]]
    step = state['step'] if 'step' in state else 1
    step = float(step)  # torch >= 1.12 keeps it as a tensor
    warmup_steps = int(round(arguments['warmup_epochs'] * arguments['num_examples'] // arguments['batch_size']))
    global_batch = arguments['world_size'] * arguments['batch_size']
    scaling = arguments['learning_rate_scaling']
    if scaling == 'linear':
        peak = arguments['base_learning_rate'] * global_batch / 256.
    elif scaling == 'sqrt':
        peak = arguments['base_learning_rate'] * math.sqrt(global_batch)
    else:
        raise ValueError('Unknown learning rate scaling {}'.format(scaling))
    if step < warmup_steps:
        lr = step / warmup_steps * peak
    else:
        total = _get_train_steps(arguments['num_examples'], arguments['train_epochs'], arguments['batch_size'])
        lr = _cosine_decay(peak, step - warmup_steps, total - warmup_steps)
    for group in opt.param_groups:
        group['lr'] = lr
    return lr


class LARC:
    """Layer-wise adaptive rate clipping around any optimiser -- what ``apex.parallel.LARC`` does in the
    reference's 'lars' branch (Model_Util.py:80-83), restated in plain PyTorch because apex is not part
    of this image.  Before every ``step()`` each parameter's gradient is rescaled by
    ``trust * ||w|| / (||g|| + wd ||w|| + eps)`` (clipped so that the effective rate never exceeds the
    group's lr when ``clip``), with the weight decay folded into the gradient."""

    def __init__(self, optimizer, trust_coefficient=0.02, clip=True, eps=1e-8):
        self.optim = optimizer
        self.trust_coefficient = trust_coefficient
        self.clip = clip
        self.eps = eps

    def __getattr__(self, name):  # state, param_groups, zero_grad, state_dict, ...
        return getattr(self.optim, name)

    @torch.no_grad()
    def step(self):
        saved = []
        for group in self.optim.param_groups:
            wd = group.get('weight_decay', 0)
            saved.append(wd)
            group['weight_decay'] = 0
            for p in group['params']:
                if p.grad is None:
                    continue
                pn, gn = torch.norm(p), torch.norm(p.grad)
                ratio = self.trust_coefficient * pn / (gn + pn * wd + self.eps)
                if self.clip:
                    ratio = torch.clamp(ratio / group['lr'], max=1.0)
                # apex leaves parameters with a zero norm (or zero gradient) untouched
                ratio = torch.where((pn > 0) & (gn > 0), ratio, torch.ones_like(ratio))
                p.grad.add_(p, alpha=wd).mul_(ratio)
        self.optim.step()
        for group, wd in zip(self.optim.param_groups, saved):
            group['weight_decay'] = wd


def get_optimizer(model, args):
    """Model_Util.py:68-88: 'sgd' (momentum, weight decay), 'adam', 'lars' (Adam inside LARC)."""
    if args.optimizer == 'sgd':
        return optim.SGD(model.parameters(), args.lr, momentum=args.momentum, weight_decay=args.weight_decay)
    if args.optimizer == 'adam':
        return optim.Adam(model.parameters(), args.lr)
    if args.optimizer == 'lars':
        return LARC(optim.Adam(model.parameters(), args.lr))
    raise ValueError('Unknown optimizer {}'.format(args.optimizer))
