"""One training step of ``convnextv2_*-unet`` with one or more mono-temporal encoders (BASELINE.json configs[4]: AERIAL_RGBI
4 ch + DEM_ELEV 1 ch; tasks_module.py:133-167 + :377-391): training-mode forward (BatchNorm on batch statistics), weighted
cross entropy, the backward through head, decoder, FusionHandler (flair_model.py:473-547) and encoders, and AdamW.

Everything numeric runs on the CUDA kernels behind the C ABI (engine/convnext_train.py, csrc/backward_ops.cu,
csrc/training_ops.cu, the tcgen05 GEMM); torch supplies memory, views, concatenation / slicing copies and dtype casts.
History of the step at configs[4] size on one B200: 235 ms (round 1) -> 182 -> 97 -> 81 -> 70 -> 67 ms, then 70.6 ms with the
fp16 forward that lifted the gradient cosine against fp32 autograd from 0.990 to 0.9987 (DESIGN.md sections 2b and 8).
BatchNorm running statistics are updated in place like nn.BatchNorm2d does (momentum 0.1, unbiased variance); under
torch.distributed the gradients are averaged DDP-style (trainers.py:81-91): every finished group of the backward -- decoder,
fusion convolutions, then each encoder stage, deepest first -- is copied into its (contiguous) range of the flat gradient
arena and that range's NCCL all-reduce starts at once on a side stream, so only the last, smallest bucket (an encoder's stem)
is exposed.

``cuda_graph=True``: the step is launch-bound on the host (about 4000 kernel launches), so after one eager step the whole
step -- rebuilding the 16-bit weight copies, forward, backward, the bucketed
all-reduces, AdamW with its step counter on the device -- is captured once into a CUDA graph and replayed; inputs are copied
into static buffers, loss and predictions come back in static buffers.  Under torch.distributed the capture is a chain of
graphs cut at the gradient buckets, with the NCCL calls launched eagerly between the segments (_capture)."""
from typing import Dict, List

import torch

from .. import native as nv
from ..flair_hub.tasks.module_setup import WeightedCrossEntropy
from ..flair_hub.tasks.tasks_module import AdamW
from .convnext_train import ACT, ConvNeXtV2EncoderTrain, UnetDecoderTrain, _to_act, _to_bf16


class ConvNeXtUNetTrainer:
    def __init__(self, state: Dict[str, torch.Tensor], depths, dims, modalities: List[str], task: str,
                 class_weight: torch.Tensor, task_weight: float = 1.0, lr: float = 5e-5, weight_decay: float = 0.01,
                 betas=(0.9, 0.999), cuda_graph: bool = False, mod_dropout: bool = False):
        """state: the model's parameters by state_dict name (fp32, CUDA); they become views into the optimizer's arena.
        ``mod_dropout``: the reference's training-time modality dropout (flair_model.py:406-408, active with more than one
        encoder): a dropped modality's feature maps are replaced by fresh noise, its encoder gets no gradient and -- like
        torch.optim.AdamW with ``grad is None`` -- is left untouched by the update.  The random decisions change the work
        from step to step, so such a trainer always steps eagerly."""
        self.mod_dropout = bool(mod_dropout)
        self.last_dropped: List[str] = []
        self.cuda_graph = bool(cuda_graph) and not self.mod_dropout
        self._graph, self._static, self._graph_out, self._stale, self._capturing = None, None, None, False, False
        self._segments, self._segmented, self._pool, self._cap_graph, self._graph_reduced = [], False, None, None, []
        self.depths, self.dims, self.mods, self.task, self.task_weight = depths, dims, list(modalities), task, task_weight
        self.names = [k for k in state if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
        self.params = {k: state[k] for k in self.names}
        self.buffers = {k: state[k] for k in state if k not in self.params}       # BatchNorm running statistics
        self.opt = AdamW([self.params[k] for k in self.names], lr=lr, weight_decay=weight_decay, betas=betas)
        self.criterion = WeightedCrossEntropy(class_weight)
        self._slot = {}                                           # parameter name -> (offset, numel) in the arena
        off = 0
        for name in self.names:
            k = self.params[name].numel()
            self._slot[name] = (off, k)
            off += k
        self.names_index = {n: i for i, n in enumerate(self.names)}
        self._comm_stream = None
        self.last_targets = None
        self._arena16 = None
        self._works, self._reduced, self._leftover, self._filled, self._overlap = [], [], [], set(), False
        self._last_ar_ms, self._ar_events = 0.0, None
        self._build()

    def _build(self) -> None:
        """Engines hold 16-bit / repacked copies of the weights: rebuilt after every optimizer step."""
        p = self.params
        # the whole fp32 arena to 16 bits in ONE launch per format (forward format for the forward GEMMs, bf16 for the data
        # gradients); the Linear weights of the blocks (the bulk of the parameters) are views into these copies, the few
        # tensors that need another layout (depthwise taps, convolutions) are still repacked one by one
        n4 = self.opt.arena.numel() // 4 * 4         # the cast kernel moves float4s; nothing that is used lives in the tail
        if self._arena16 is None:
            self._arena16 = {fmt: torch.empty(self.opt.arena.shape, dtype=dt, device=self.opt.arena.device)
                             for fmt, dt in (("act", ACT), ("bf16", torch.bfloat16))}
        p16 = {"act": {}, "bf16": {}}
        for fmt, arena16 in self._arena16.items():
            if fmt == "bf16" and ACT == torch.bfloat16:
                arena16 = self._arena16["act"]                       # one format: one copy
            else:
                nv.cast_f32_bf16(self.opt.arena[:n4], arena16[:n4])
            for name in self.names:
                if name.endswith(("mlp.fc1.weight", "mlp.fc2.weight")):
                    off, k = self._slot[name]
                    if off % 8 == 0 and off + k <= n4:   # 16-byte aligned rows for the TMA descriptor
                        p16[fmt][name] = arena16[off:off + k].view(p[name].shape)
        self.enc = {}
        for m in self.mods:
            pre = f"encoders.{m}.seg_model.model."
            self.enc[m] = ConvNeXtV2EncoderTrain({k[len(pre):]: v for k, v in p.items() if k.startswith(pre)}, self.depths,
                                                 self.dims, params16={fmt: {k[len(pre):]: v for k, v in d16.items() if k.startswith(pre)}
                                                                      for fmt, d16 in p16.items()})
        pre = f"main_decoders.{self.task}.seg_model."
        both = {**p, **self.buffers}
        self.dec = UnetDecoderTrain({k[len(pre):]: v for k, v in both.items() if k.startswith(pre)})
        self.fuse = None
        if len(self.mods) > 1:
            # (forward-format weight, bf16 weight for the data gradient, bias)
            self.fuse = [(p[f"fusion_handler.conv_f.{i}.weight"].detach().reshape(c, -1).to(ACT).contiguous(),
                          p[f"fusion_handler.conv_f.{i}.weight"].detach().reshape(c, -1).to(torch.bfloat16).contiguous(),
                          p[f"fusion_handler.conv_f.{i}.bias"].detach().float().contiguous()) for i, c in enumerate(self.dims)]

    # ------------------------------------------------------------------ gradient arena + bucketed all-reduce
    @property
    def last_allreduce_ms(self) -> float:
        """EXPOSED part of the last step's gradient exchange: what the compute stream waited for after the backward.  Graph
        replays time it with a pair of events read back here, once they have completed (no host stall inside the step)."""
        if self._ar_events is not None and self._ar_events[1].query():
            self._last_ar_ms = self._ar_events[0].elapsed_time(self._ar_events[1])
            self._ar_events = None
        return self._last_ar_ms

    @last_allreduce_ms.setter
    def last_allreduce_ms(self, v: float) -> None:
        self._last_ar_ms, self._ar_events = float(v), None

    def _distributed(self) -> bool:
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def _emit(self, group: Dict[str, torch.Tensor]) -> None:
        """Copies a finished group of gradients into the arena; under torch.distributed starts the all-reduce (average) of
        the arena range the group covers on the communication stream.  Groups cover contiguous ranges (parameters of a
        module are adjacent in state_dict order); a group that does not is left to the final catch-all reduction."""
        if not group:
            return
        lo, hi, tot = None, None, 0
        for name, g in group.items():
            off, k = self._slot[name]
            self.opt.grads[self.names_index[name]].copy_(g.reshape(self.opt.grads[self.names_index[name]].shape))
            lo = off if lo is None else min(lo, off)
            hi = off + k if hi is None else max(hi, off + k)
            tot += k
            self._filled.add(name)
        if self._capturing:
            # graph capture: NCCL stays OUT of the graphs.  A contiguous bucket ends the current graph segment; the replay
            # launches its all-reduce eagerly on the side stream between two segment replays (see _replay)
            if self._segmented and tot == hi - lo:
                self._seg_end(("reduce", lo, hi))
                self._seg_begin()
                self._reduced.append((lo, hi))
            return
        if not self._overlap:
            return
        if tot != hi - lo:
            self._leftover.append((lo, hi))
            return
        self._launch_allreduce(lo, hi)
        self._reduced.append((lo, hi))

    def _launch_allreduce(self, lo: int, hi: int) -> None:
        """all-reduce (average) of grad[lo:hi] on the communication stream, ordered after everything already on the current one"""
        import torch.distributed as dist
        if self._comm_stream is None:
            self._comm_stream = torch.cuda.Stream(device=self.opt.grad.device)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        with torch.cuda.stream(self._comm_stream):
            self._comm_stream.wait_event(ev)
            self._works.append(dist.all_reduce(self.opt.grad[lo:hi], op=dist.ReduceOp.AVG, async_op=True))

    def forward_backward(self, batch: Dict[str, torch.Tensor], into_arena: bool = False):
        """-> (loss 0-d tensor, preds int32 (B,H,W), {parameter name: gradient}).  With ``into_arena`` the gradients are also
        written to the optimizer's gradient arena group by group as the backward produces them (and all-reduced, see _emit)."""
        emit = self._emit if into_arena else None
        if self._stale and not self._capturing:      # graph replays rebuild the weight copies inside the graph only
            self._build()
            self._stale = False
        dropped = {}
        if self.mod_dropout and into_arena and len(self.mods) > 1:
            from ..flair_hub.models.flair_model import draw_modality_dropout
            shapes = {}
            for m in self.mods:                      # smp's feature list for these encoders: x, 0-channel dummy, f4 .. f32
                B, cin, P, _ = batch[m].shape
                shapes[m] = [(B, cin, P, P), (B, 0, P // 2, P // 2)] + [(B, c, P // (4 << i), P // (4 << i))
                                                                         for i, c in enumerate(self.dims)]
            dropped = draw_modality_dropout(shapes, batch[self.mods[0]].device)
        self.last_dropped = list(dropped)
        feats = {}
        for m in self.mods:
            if m in dropped:                         # the encoder's output is discarded: its forward is skipped altogether
                feats[m] = [t.permute(0, 2, 3, 1).contiguous() for t in dropped[m][2:]]
            else:
                feats[m] = self.enc[m].forward(batch[m])
        cats = None
        if self.fuse is None:
            fused = feats[self.mods[0]]
        else:
            fused, cats = [], []
            for i, (w, _, b) in enumerate(self.fuse):
                cat = torch.cat([_to_act(feats[m][i]) for m in self.mods], dim=-1).contiguous()       # [B,h,h,sum C]
                B, h, _, ct = cat.shape
                cats.append(cat)
                fused.append(nv.gemm_bf16(cat.view(-1, ct), w, nv.EPI_F32, bias=b).view(B, h, h, -1))
        logits = self.dec.forward(fused)
        targets = batch[self.task]
        targets = nv.onehot_argmax(targets) if targets.dim() == 4 else targets.to(torch.int32)
        self.last_targets = targets                  # int32 (B,H,W): what the metrics compare the predictions with
        loss, preds = self.criterion(logits, targets, task_weight=self.task_weight, want_preds=True)
        dfused, grads = self.dec.backward(self.criterion.backward())
        grads = {f"main_decoders.{self.task}.seg_model.{k}": v for k, v in grads.items()}
        if emit:
            emit(grads)
        if self.fuse is None:
            dfeats = {self.mods[0]: dfused}
        else:
            dfeats = {m: [] for m in self.mods}
            for i, (_, w, b) in enumerate(self.fuse):
                cat = cats[i]
                B, h, _, ct = cat.shape
                dcat, dw, db = nv.linear_backward(_to_bf16(dfused[i].reshape(-1, self.dims[i])), cat.view(-1, ct), w)
                grads[f"fusion_handler.conv_f.{i}.weight"] = dw.view(self.dims[i], ct, 1, 1)
                grads[f"fusion_handler.conv_f.{i}.bias"] = db
                off = 0
                for m in self.mods:
                    c = feats[m][i].shape[-1]
                    dfeats[m].append(dcat[:, off:off + c].float().reshape(B, h, h, c).contiguous())
                    off += c
            if emit:
                emit({k: v for k, v in grads.items() if k.startswith("fusion_handler.")})
        for m in self.mods:
            if m in dropped:                         # noise has no producer: no gradient for this encoder
                continue
            pre = f"encoders.{m}.seg_model.model."
            g = self.enc[m].backward(dfeats[m], emit=(lambda part, pre=pre: emit({pre + k: v for k, v in part.items()})) if emit else None)
            grads.update({pre + k: v for k, v in g.items()})
        return loss, preds, grads

    def step(self, batch: Dict[str, torch.Tensor]):
        """forward + backward + (overlapped) gradient all-reduce + AdamW update; -> (loss before the update, preds).
        With ``cuda_graph`` the first call runs eagerly (it sizes the library's scratch buffers and initialises NCCL), the
        second captures, and every call from the second on replays the graph; the returned loss / preds are then the graph's
        static output buffers (overwritten by the next step).  A batch of another shape falls back to the eager step."""
        if not self.cuda_graph or self.opt.step_count == 0:
            return self._step_eager(batch)
        if self._graph is not None and not self._same_shapes(batch):
            return self._step_eager(batch)
        if self._graph is None:
            self._capture(batch)
        for k, v in self._static.items():
            v.copy_(batch[k], non_blocking=True)
        self._replay()
        self.opt.note_device_step()
        self._stale = True
        return self._graph_out

    def set_lr(self, lr: float) -> None:
        """New learning rate from a scheduler (flair_hub/tasks/schedulers.py).  The captured graphs carry the rate as a kernel
        argument, so a change drops them and the next step captures again: fine for a plateau schedule that moves a few times
        per run; a per-step schedule should run the trainer with ``cuda_graph=False``."""
        lr = float(lr)
        if lr != self.opt.lr:
            self.opt.lr = lr
            self._graph, self._segments, self._graph_out = None, [], None

    def _same_shapes(self, batch) -> bool:
        return all(k in batch and batch[k].shape == v.shape and batch[k].dtype == v.dtype for k, v in self._static.items())

    def _capture(self, batch) -> None:
        """One graph in a single process.  Under torch.distributed (or FZ_TRAIN_SEGMENTS=1, which exercises the same code on
        one GPU) the step is captured as a CHAIN of graphs sharing one memory pool, cut wherever a gradient bucket is complete
        and once more in front of the optimizer: NCCL is never captured (a capture containing the side-stream all-reduces did
        not come back on 2 GPUs); the replay launches each bucket's all-reduce eagerly between two segments, so the exchange
        still overlaps the rest of the backward."""
        import os
        keys = list(self.mods) + [self.task]
        self._static = {k: batch[k].detach().clone() for k in keys}
        self._segmented = self._distributed() or os.environ.get("FZ_TRAIN_SEGMENTS", "0") == "1"
        torch.cuda.synchronize()
        torch.cuda.empty_cache()                     # the eager step's cached blocks: the graphs get their own pool
        self._segments, self._reduced = [], []
        self._pool = torch.cuda.graph_pool_handle()
        cap = torch.cuda.Stream(device=self.opt.grad.device)
        cap.wait_stream(torch.cuda.current_stream())
        self._capturing = True
        try:
            with torch.cuda.stream(cap):
                self._seg_begin()
                self._build()                        # inside the graph: replays refresh the 16-bit weight copies themselves
                self._graph_out = self._step_body(self._static, timed=False, on_device_counter=True)
                self._seg_end(None)
        finally:
            self._capturing = False
        torch.cuda.current_stream().wait_stream(cap)
        self._graph = self._segments[0][0]
        self._graph_reduced = sorted(self._reduced)

    def _seg_begin(self) -> None:
        self._cap_graph = torch.cuda.CUDAGraph()
        self._cap_graph.capture_begin(pool=self._pool, capture_error_mode="thread_local")     # NCCL's watchdog thread keeps polling

    def _seg_end(self, action) -> None:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")          # two cuts in a row leave an empty segment: harmless, torch warns
            self._cap_graph.capture_end()
        self._segments.append((self._cap_graph, action))
        self._cap_graph = None

    def _replay(self) -> None:
        import torch.distributed as dist
        live = self._distributed()
        self._works = []
        for graph, action in self._segments:
            graph.replay()
            if action is None or not live:
                continue
            if action[0] == "reduce":
                self._launch_allreduce(action[1], action[2])
            elif action[0] == "finish":              # everything the buckets did not cover, then wait for the buckets
                cur = torch.cuda.current_stream()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cur)
                pos, n = 0, self.opt.grad.numel()
                for lo, hi in self._graph_reduced + [(n, n)]:
                    if lo > pos:
                        dist.all_reduce(self.opt.grad[pos:lo], op=dist.ReduceOp.AVG)
                    pos = max(pos, hi)
                for w in self._works:
                    w.wait()
                if self._comm_stream is not None:
                    cur.wait_stream(self._comm_stream)
                e1.record(cur)
                self._ar_events = (e0, e1)

    def _step_eager(self, batch):
        if self._stale:
            self._build()
            self._stale = False
        out = self._step_body(batch, timed=True, on_device_counter=False)
        self._build()
        return out

    def _step_body(self, batch, timed: bool, on_device_counter: bool):
        self._filled, self._reduced, self._leftover, self._works = set(), [], [], []
        # modality dropout is drawn per rank: which buckets exist would differ between ranks, so the exchange is then ONE
        # all-reduce of the whole arena after the backward
        self._overlap = self._distributed() and not self.mod_dropout
        loss, preds, grads = self.forward_backward(batch, into_arena=True)
        # parameters the forward never touches (fusion_handler.conv_f with a single modality) have no gradient: torch's AdamW
        # leaves them alone (no weight decay either), so they are put back after the fused update
        dropped_everywhere = self.last_dropped
        if self.mod_dropout and self._distributed():
            # DDP semantics (find_unused_parameters): an encoder unused on SOME ranks still gets the averaged gradient of the
            # others; only one that every rank dropped has no gradient at all
            import torch.distributed as dist
            used = torch.tensor([0.0 if m in self.last_dropped else 1.0 for m in self.mods], device=self.opt.grad.device)
            dist.all_reduce(used)
            dropped_everywhere = [m for m, u in zip(self.mods, used.tolist()) if u == 0]
        skip = []
        for m in dropped_everywhere:                 # a dropped encoder: one contiguous arena range the update leaves alone
            pre = f"encoders.{m}."
            offs = [self._slot[n] for n in self.names if n.startswith(pre)]
            skip.append((min(o for o, _ in offs), max(o + k for o, k in offs)))
        unused = {n: self.params[n].clone() for n in self.names
                  if n not in self._filled and not any(lo <= self._slot[n][0] < hi for lo, hi in skip)}
        for n in self.names:
            if n not in self._filled:
                self.opt.grads[self.names_index[n]].zero_()          # so that the averaged gradient of a skipped range is 0 too
        if self.mod_dropout and self._distributed():
            self.allreduce_gradients()
        else:
            self._finish_allreduce(timed)
        if on_device_counter:
            self.opt.step_on_device_counter()
        else:
            self.opt.step(skip=skip)
        for n, v in unused.items():
            self.params[n].copy_(v)
        return loss, preds

    def _finish_allreduce(self, timed: bool = True) -> None:
        """Waits for the buckets started during the backward and reduces whatever they did not cover (parameters without a
        gradient, non-contiguous groups).  ``last_allreduce_ms`` = the time the compute stream had to wait, i.e. the EXPOSED
        part of the gradient exchange (``timed=False``, under graph capture: not measured, the last eager value is kept)."""
        if self._capturing:
            if self._segmented:
                self._seg_end(("finish",))
                self._seg_begin()
            return
        if timed:
            self.last_allreduce_ms = 0.0
        if not self._overlap:
            return
        import torch.distributed as dist
        cur = torch.cuda.current_stream()
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
        for w in self._works:
            w.wait()                                  # the compute stream waits for NCCL's stream
        if self._comm_stream is not None:
            cur.wait_stream(self._comm_stream)        # explicit join (a forked stream must rejoin before a capture ends)
        covered = sorted(self._reduced)
        pos, n = 0, self.opt.grad.numel()
        for lo, hi in covered + [(n, n)]:
            if lo > pos:                              # a gap no bucket covered: reduce it now (tiny: unused parameters)
                dist.all_reduce(self.opt.grad[pos:lo], op=dist.ReduceOp.AVG)
            pos = max(pos, hi)
        if timed:
            e1.record(cur)
            e1.synchronize()
            self.last_allreduce_ms = e0.elapsed_time(e1)

    def allreduce_gradients(self) -> float:
        """The un-overlapped exchange (round 1; kept for A/B): ONE blocking NCCL all-reduce of the whole gradient arena
        (183 M fp32 for configs[4]).  Returns the milliseconds it took on this rank (0.0 when not distributed).  Decoder
        BatchNorm statistics stay per-GPU (the reference uses no SyncBN)."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return 0.0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        dist.all_reduce(self.opt.grad, op=dist.ReduceOp.AVG)
        e1.record()
        e1.synchronize()
        self.last_allreduce_ms = e0.elapsed_time(e1)
        return self.last_allreduce_ms
