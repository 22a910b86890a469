"""The ``Model`` protocol of the reference (realpdebench/model/model.py:4-26).

If the reference package is importable its own ``Model`` class is used as the
base, so ``isinstance(m, realpdebench.model.model.Model)`` holds for engine
models; otherwise an identical mirror is defined here.
"""
import torch
import torch.nn as nn


def _reference_model_base():
    try:
        from realpdebench.model.model import Model as RefModel  # noqa: WPS433
        return RefModel
    except Exception:
        return None


_Ref = _reference_model_base()

if _Ref is not None:
    Model = _Ref
else:
    class Model(nn.Module):
        def __init__(self):
            super().__init__()

        def forward(self, x):
            raise Exception(NotImplementedError)

        def train_loss(self, input, target):
            raise Exception(NotImplementedError)

        def load_checkpoint(self, checkpoint_path, device):
            # model.py:14-26
            checkpoint = torch.load(checkpoint_path, map_location=device)
            self.load_state_dict(checkpoint['model_state_dict'])
            return {
                'all_train_losses': checkpoint['train_losses'],
                'all_val_losses': checkpoint['val_losses'],
                'iteration': checkpoint['iteration'],
                'best_iteration': checkpoint['best_iteration'],
                'best_val_loss': checkpoint['best_val_loss'],
            }
