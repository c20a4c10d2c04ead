"""``FLAIRLosses``: drop-in for flair_hub/tasks/module_setup.py:119-200 on the training step's loss side, and
the two factories in front of it: ``build_segmentation_module`` (:47-83) and ``get_input_img_sizes`` (:87-117).  The data
module (``build_data_module``, :12-43: the FLAIR-HUB patch datasets and augmentations) is outside the scope of SURVEY 8.

The reference builds one ``nn.CrossEntropyLoss(weight=w)`` per task (``_create_task_loss``, :155) with
``w = [value_weights.default] * n_classes`` overridden by ``value_weights.default_exceptions`` (:180-196), plus auxiliary
losses per modality (:141-146, out of scope here like the auxiliary decoders).  Here each criterion is a
``WeightedCrossEntropy`` whose forward and gradient are the CUDA kernels of csrc/training_ops.cu behind the C ABI."""
from typing import Dict, Optional

import torch

from ... import native as nv


class WeightedCrossEntropy:
    """``nn.CrossEntropyLoss(weight=w)`` ('mean' reduction) on (B,C,H,W) fp32 logits and (B,H,W) class indices."""

    def __init__(self, weight: torch.Tensor):
        self.weight = weight.float()
        self._saved = None

    def to(self, device):
        self.weight = self.weight.to(device)
        return self

    def __call__(self, logits: torch.Tensor, targets: torch.Tensor, task_weight: float = 1.0, want_preds: bool = False):
        """-> loss (0-d tensor, = task_weight * CE); keeps what ``backward`` needs.  With ``want_preds`` also returns
        argmax(softmax(logits)) (tasks_module.py:159)."""
        if not logits.is_cuda:
            raise nv.NativeError("WeightedCrossEntropy runs on CUDA only (no CPU fallback)")
        logits = logits.float().contiguous()
        targets = targets.to(torch.int32).contiguous()
        w = self.weight.to(logits.device)
        out, lse, preds = nv.ce_loss_forward(logits, targets, w, task_weight, want_preds=want_preds)
        self._saved = (logits, targets, w, float(task_weight), lse, out)
        return (out[0], preds) if want_preds else out[0]

    def backward(self, grad_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """d loss / d logits of the last call, (B,C,H,W) fp32."""
        if self._saved is None:
            raise RuntimeError("backward() before a forward call")
        logits, targets, w, tw, lse, loss_out = self._saved
        return nv.ce_loss_backward(logits, targets, w, tw, lse, loss_out, grad_scale, out=out)


class FLAIRLosses:
    def __init__(self, config: Dict[str, dict]) -> None:
        self.config = config
        self.default_weights: Dict[str, torch.Tensor] = {}
        self.losses: Dict[str, WeightedCrossEntropy] = self._build_losses()

    def _build_losses(self) -> Dict[str, WeightedCrossEntropy]:
        losses = {}
        for task in self.config['labels']:
            task_config = self.config['labels_configs'][task]
            default_w = self._compute_default_weights(task_config)
            self.default_weights[task] = default_w
            losses[task] = WeightedCrossEntropy(default_w)
            for modality, aux_active in (self.config.get('modalities', {}).get('aux_loss', {}) or {}).items():
                if aux_active and self.config['modalities']['inputs'].get(modality, False):
                    raise NotImplementedError("auxiliary losses belong to the auxiliary decoders, which are outside the "
                                              "zonal hot path (SURVEY.md section 8)")
        return losses

    @staticmethod
    def _compute_default_weights(task_config: Dict[str, dict]) -> torch.Tensor:
        """module_setup.py:180-196."""
        vw = task_config['value_weights']
        w = torch.tensor([float(vw['default'])] * len(task_config['value_name']), dtype=torch.float32)
        for key, value in (vw.get('default_exceptions') or {}).items():
            w[int(key)] = float(value)
        return w

    def get_losses(self) -> Dict[str, WeightedCrossEntropy]:
        return self.losses


def build_segmentation_module(config: dict, in_img_sizes, stage: str = 'train'):
    """module_setup.py:47-83: the model for ``in_img_sizes`` wrapped in a ``SegmentationTask`` -- with the losses for
    'train', without for 'predict'."""
    assert stage in ['train', 'predict'], "stage must be either 'train' or 'predict'"
    from ..models.flair_model import FLAIR_HUB_Model
    from .tasks_module import SegmentationTask
    model = FLAIR_HUB_Model(config, in_img_sizes)
    if stage == 'train':
        return SegmentationTask(model=model, config=config, criterion=FLAIRLosses(config).get_losses())
    return SegmentationTask(model=model, config=config)


def get_input_img_sizes(config: dict, dm, stage: str = "fit") -> Dict[str, int]:
    """module_setup.py:87-117: the last dimension of each active input modality in the first batch of the data module's
    train / predict loader (any object with ``setup``, ``train_dataloader``, ``predict_dataloader``)."""
    assert stage in {"fit", "predict"}, f"Unsupported stage '{stage}'"
    dm.setup(stage)
    dataloader = dm.train_dataloader() if stage == "fit" else dm.predict_dataloader()
    monkeybatch = next(iter(dataloader))
    return {modality: monkeybatch[modality][0].shape[-1] for modality, is_input in config['modalities']['inputs'].items()
            if is_input and modality in monkeybatch}
