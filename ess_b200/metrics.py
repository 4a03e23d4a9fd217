"""Segmentation metrics with the confusion matrix built on the device (reference surface:
evaluation/metrics.py -- `MetricsSemseg.update_batch / get_metrics_summary / reset`, confusion[y, y_hat],
per-class IoU in percent, mean IoU, pixel accuracy).

Unlike the reference, which bincounts on the device and then copies the K x K matrix to the host after
every batch, the running matrix stays on the GPU as int64 and is only brought to the host when a summary
is requested -- one D2H copy per evaluation instead of one per batch."""
import torch

from . import ops


def semseg_compute_confusion(y_hat_lbl, y_lbl, num_classes, ignore_label, out=None):
    """int64 [K, K] confusion counts (rows = ground truth, columns = prediction) of non-ignored pixels."""
    if not (torch.is_tensor(y_hat_lbl) and torch.is_tensor(y_lbl)):
        raise TypeError('Inputs must be torch tensors')
    if y_hat_lbl.device != y_lbl.device:
        raise ValueError('Input tensors have different device placement')
    with ops.on_device_of(y_hat_lbl):
        pred = y_hat_lbl.reshape(-1).long().contiguous()
        gt = y_lbl.reshape(-1).long().contiguous()
        if pred.numel() != gt.numel():
            raise ValueError('prediction / label size mismatch')
        return ops.confusion_labels(pred, gt, num_classes, ignore_label, conf=out)


def _iou_acc(conf):
    c = conf.to(torch.float64)
    tp = torch.diagonal(c)
    union = c.sum(0) + c.sum(1) - tp
    iou = 100.0 * tp / union.clamp(min=1e-12)
    acc = 100.0 * tp.sum() / c.sum().clamp(min=1e-12)
    return iou, acc


def semseg_accum_confusion_to_iou(confusion_accum):
    iou, _ = _iou_acc(confusion_accum)
    return iou.mean(), iou


def semseg_accum_confusion_to_acc(confusion_accum):
    return _iou_acc(confusion_accum)[1]


class MetricsSemseg:
    def __init__(self, num_classes, ignore_label, class_names):
        self.num_classes, self.ignore_label, self.class_names = num_classes, ignore_label, class_names
        self._running = None        # int64 [K, K] on the device

    def reset(self):
        self._running = None

    @property
    def metrics_acc(self):
        """Host copy of the accumulated confusion matrix (the reference keeps this attribute on the CPU)."""
        return None if self._running is None else self._running.cpu()

    def update_batch(self, y_hat_lbl, y_lbl):
        with torch.no_grad():
            if self._running is None:
                self._running = torch.zeros((self.num_classes, self.num_classes), device=y_lbl.device,
                                            dtype=torch.int64)
            semseg_compute_confusion(y_hat_lbl, y_lbl, self.num_classes, self.ignore_label, out=self._running)

    def get_metrics_summary(self):
        cm = self.metrics_acc
        iou, acc = _iou_acc(cm)
        summary = {name: iou[i] for i, name in enumerate(self.class_names)}
        summary.update(mean_iou=iou.mean(), acc=acc, cm=cm)
        return summary
