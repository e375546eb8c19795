"""Post-backbone inference wiring of `GeneralizedRCNN.inference` (lib/modeling/meta_arch/rcnn.py:92-144):

    RPN.call -> find_top_rpn_proposals           (lib/modeling/proposal_generator/rpn.py:143-195)
    StandardROIHeads._forward_box  -> box pooler + fast_rcnn_inference   (roi_heads.py:545-577)
    StandardROIHeads._forward_mask -> mask pooler on the detections      (roi_heads.py:579-605)

with the dense conv/FC layers (RPN head, box head, mask head) outside: their outputs are inputs here.
Only reference-facing operators of this package are called; everything runs on torch's current stream
with no host synchronisation (static instance grids replace the reference's tf.where compaction: every
slot of the zero-padded dense outputs is pooled, `is_valid` tells which rows are live).

`run_host` is the host-buffer entry point (the analogue of `sess.run(fetches, feed_dict)`): inputs are host
tensors, outputs are pinned host tensors.  Images are independent, so the batch is cut into chunks that are
software-pipelined over two CUDA streams: the upload of chunk i+1, the kernels of chunk i and the download
of chunk i-1 overlap (PCIe is full duplex).
"""
import numpy as np
import torch

from .modeling import Box2BoxTransform, ROIPooler, RPNOutputs, fast_rcnn_inference
from .structures import BoxList, ImageList, SparseBoxList

PER_IMAGE_KEYS = ("logits", "deltas", "feats", "shapes")   # tensors (or lists of) with a leading image dim
PER_ROI_KEYS = ("scores", "cls_deltas")                      # [N * rois_per_image, ...], image-major


class MaskRCNNPostBackbone(object):
    def __init__(self, rois_per_image=1000, dets_per_image=100, pre_nms_topk=2000, rpn_nms_thresh=0.7,
                 min_box_side_len=0.0, score_thresh=0.05, nms_thresh=0.5, nms_cls_agnostic=False,
                 scales=(1 / 4., 1 / 8., 1 / 16., 1 / 32.), box_resolution=7, mask_resolution=14, sampling_ratio=0,
                 pooler_type="ROIAlignV2", rpn_weights=(1.0, 1.0, 1.0, 1.0), box_weights=(10.0, 10.0, 5.0, 5.0)):
        self.R, self.D = int(rois_per_image), int(dets_per_image)
        self.pre, self.rpn_thr, self.min_len = int(pre_nms_topk), float(rpn_nms_thresh), float(min_box_side_len)
        self.score_thr, self.nms_thr, self.agnostic = float(score_thresh), float(nms_thresh), bool(nms_cls_agnostic)
        self.rpn_tf = Box2BoxTransform(rpn_weights)
        self.box_tf = Box2BoxTransform(box_weights)
        self.box_pooler = ROIPooler(box_resolution, list(scales), sampling_ratio, pooler_type)
        self.mask_pooler = ROIPooler(mask_resolution, list(scales), sampling_ratio, pooler_type)
        self._grids = {}
        self._streams = {}
        self._host_out = {}

    # ------------------------------------------------------------------ helpers
    def _grid(self, n, per, dev):
        key = (n, per, str(dev))
        g = self._grids.get(key)
        if g is None:
            img = np.repeat(np.arange(n, dtype=np.int64), per)
            slot = np.tile(np.arange(per, dtype=np.int64), n)
            g = torch.from_numpy(np.stack([img, slot], 1)).to(dev)
            self._grids[key] = g
        return g

    # ------------------------------------------------------------------ device-resident step
    def __call__(self, x, events=None):
        """x: dict of DEVICE tensors -- logits L x [n,HWA], deltas L x [n,HWA,4], anchors L x [HWA,4],
        feats 4 x [n,H,W,C], shapes [n,2] int32, scores [n*R,K+1], cls_deltas [n*R,K*4].
        Returns dict(proposals BoxList, box_feats, dets BoxList, mask_feats)."""
        def mark(i):
            if events is not None:
                events[i].record()
        n = x["shapes"].shape[0]
        dev = x["shapes"].device
        mark(0)
        outs = RPNOutputs(self.rpn_tf, ImageList(None, x["shapes"]), x["logits"], x["deltas"], x["anchors"])
        props = outs.find_top_proposals(self.rpn_thr, self.pre, self.R, self.min_len)
        mark(1)
        inst = SparseBoxList(self._grid(n, self.R, dev), BoxList(props.boxes.reshape(-1, 4)), (n, self.R))
        inst.set_tracking("image_shape", x["shapes"])
        box_feats = self.box_pooler(x["feats"], inst)
        mark(2)
        boxes = self.box_tf.apply_deltas(x["cls_deltas"], inst.data.boxes)
        dets, _ = fast_rcnn_inference(boxes, x["scores"], inst, self.score_thr, self.nms_thr, self.D, self.agnostic)
        mark(3)
        dinst = SparseBoxList(self._grid(n, self.D, dev), BoxList(dets.boxes.reshape(-1, 4)), (n, self.D))
        mask_feats = self.mask_pooler(x["feats"], dinst)
        mark(4)
        return dict(proposals=props, box_feats=box_feats, dets=dets, mask_feats=mask_feats)

    # ------------------------------------------------------------------ host-buffer step
    @staticmethod
    def flatten_outputs(out):
        p, d = out["proposals"], out["dets"]
        return {"proposal_boxes": p.boxes, "proposal_logits": p.get_field("objectness_logits"),
                "proposal_valid": p.get_field("is_valid"), "box_feats": out["box_feats"],
                "det_boxes": d.boxes, "det_scores": d.get_field("scores"), "det_classes": d.get_field("pred_classes"),
                "det_valid": d.get_field("is_valid"), "mask_feats": out["mask_feats"]}

    def run_host(self, x, device=None, chunk_images=2):
        """x: dict of HOST tensors (pinned for full PCIe rate).  Returns a dict of pinned host tensors
        (flatten_outputs keys) for the whole batch, byte-identical to the un-chunked device step.  The
        returned buffers are owned by this object and are overwritten by the next `run_host` call."""
        dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        N = x["shapes"].shape[0]
        R, D = self.R, self.D
        key = str(dev)
        if key not in self._streams:
            self._streams[key] = [torch.cuda.Stream(dev), torch.cuda.Stream(dev)]
        streams = self._streams[key]
        cur = torch.cuda.current_stream(dev)
        for s in streams:
            s.wait_stream(cur)
        anchors = [a.to(dev, non_blocking=True) for a in x["anchors"]]
        ev_anchor = torch.cuda.Event()
        ev_anchor.record(cur)
        host_out = None
        for ci, b in enumerate(range(0, N, chunk_images)):
            e = min(b + chunk_images, N)
            s = streams[ci % 2]
            with torch.cuda.stream(s):
                s.wait_event(ev_anchor)
                xd = {"anchors": anchors}
                for k in PER_IMAGE_KEYS:
                    v = x[k]
                    xd[k] = [t[b:e].to(dev, non_blocking=True) for t in v] if isinstance(v, (list, tuple)) \
                        else v[b:e].to(dev, non_blocking=True)
                for k in PER_ROI_KEYS:
                    xd[k] = x[k][b * R:e * R].to(dev, non_blocking=True)
                out = self.flatten_outputs(self(xd))
                if host_out is None:
                    host_out = self._host_buffers(out, N, e - b)
                for k, t in out.items():
                    per = t.shape[0] // (e - b)
                    host_out[k][b * per:e * per].copy_(t, non_blocking=True)
        for s in streams:
            cur.wait_stream(s)
        cur.synchronize()
        return host_out

    def _host_buffers(self, out, N, n_chunk):
        sig = tuple((k, tuple(t.shape[1:]), t.dtype, t.shape[0] // n_chunk) for k, t in out.items()) + (N,)
        if self._host_out.get("sig") != sig:
            bufs = {k: torch.empty((N * (t.shape[0] // n_chunk),) + tuple(t.shape[1:]), dtype=t.dtype, pin_memory=True)
                    for k, t in out.items()}
            self._host_out = {"sig": sig, "bufs": bufs}
        return self._host_out["bufs"]
