"""Host side of the ER-DDPG update (csrc/update.cu) and the replay buffer (csrc/replay.cu).

DDPGLearner owns the flat parameter / Adam buffers in the reference's parameters() order and exposes
per-tensor views keyed like the reference state_dict (ddpg_agent.py:526-560), so checkpoints interchange.
"""
import ctypes as C

import torch

from . import _lib as L

ACTOR_LAYOUT = [("emotion_weights", (3, 10)), ("input_layer.weight", (128, 5)), ("input_layer.bias", (128,)),
                ("hidden.2.weight", (128, 128)), ("hidden.2.bias", (128,)), ("hidden.5.weight", (64, 128)),
                ("hidden.5.bias", (64,)), ("hidden.7.weight", (10, 64)), ("hidden.7.bias", (10,))]
CRITIC_LAYOUT = [("net.0.weight", (256, 15)), ("net.0.bias", (256,)), ("net.1.weight", (256,)), ("net.1.bias", (256,)),
                 ("net.4.weight", (256, 256)), ("net.4.bias", (256,)), ("net.5.weight", (256,)), ("net.5.bias", (256,)),
                 ("net.8.weight", (128, 256)), ("net.8.bias", (128,)), ("net.10.weight", (1, 128)), ("net.10.bias", (1,))]


def _views(flat, layout):
    out, o = {}, 0
    for name, shape in layout:
        n = 1
        for d in shape:
            n *= d
        out[name] = flat[o:o + n].view(*shape)
        o += n
    assert o == flat.numel()
    return out


class FlatNet:
    def __init__(self, layout, total, device):
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views = _views(self.flat, layout)
        self.layout = layout

    def load(self, sd):
        for name, _ in self.layout:
            self.views[name].copy_(torch.as_tensor(sd[name]).to(self.flat.device, torch.float32))

    def state_dict(self):
        return {k: v.clone() for k, v in self.views.items()}


class ReplayBuffer:
    """[capacity][20] binary32 records on the device (layout in include/erlfill.h)."""

    def __init__(self, capacity, device="cuda"):
        self.capacity = int(capacity)
        self.device = torch.device(device)
        self.buf = torch.zeros(self.capacity, L.REC, dtype=torch.float32, device=self.device)
        self.size = 0
        self.inserted = 0       # total transitions ever inserted (ring position)

    def __len__(self):
        return self.size

    def insert(self, state, action, reward, next_state, done, emotion, scale=None, offset=None, policy="ring", cond=None):
        """R1.  policy "ring": FIFO ring (TD3/SAC deque semantics and the batched ER-DDPG loop);
        "per_argmin": the reference's PrioritizedReplayBuffer.add with its all-equal priorities — append until
        full, then always overwrite slot 0 (np.argmin of equal values).
        cond: an _lib.CondScaling for a mixed-condition batch (per-env scale / offset rows, condition id into column 19)."""
        n = state.shape[0]
        pos = None
        if policy == "per_argmin":
            free = self.capacity - self.size
            idx = [self.size + i if i < free else 0 for i in range(n)]
            pos = torch.tensor(idx, dtype=torch.int64, device=self.device)
            start = 0
        else:
            start = self.inserted % self.capacity
        with torch.cuda.device(self.device):
            if cond is not None:
                L.check(L.lib().erlfill_replay_insert_cond(C.c_int(n), L.ptr(self.buf), C.c_int64(self.capacity), C.c_int64(start),
                                                           L.ptr(pos), L.ptr(state), L.ptr(action), L.ptr(reward), L.ptr(next_state),
                                                           L.ptr(done), L.ptr(emotion), C.byref(cond), L.stream_ptr()),
                        "erlfill_replay_insert_cond")
            else:
                L.check(L.lib().erlfill_replay_insert(C.c_int(n), L.ptr(self.buf), C.c_int64(self.capacity), C.c_int64(start),
                                                      L.ptr(pos), L.ptr(state), L.ptr(action), L.ptr(reward), L.ptr(next_state),
                                                      L.ptr(done), L.ptr(emotion), L.ptr(scale), L.ptr(offset), L.stream_ptr()),
                        "erlfill_replay_insert")
        self.inserted += n
        self.size = min(self.capacity, self.size + n)

    def sample_indices(self, batch, seed, update_idx, stream_id=0, out=None):
        if out is None:
            out = torch.empty(batch, dtype=torch.int64, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_replay_sample(C.c_int(batch), C.c_int64(self.size), C.c_uint64(seed),
                                                  C.c_uint32(update_idx & 0xFFFFFFFF), C.c_uint32(stream_id), L.ptr(out),
                                                  L.stream_ptr()), "erlfill_replay_sample")
        return out

    def sample_gather(self, batch, seed, update_idx, stream_id=0, out=None, idx_out=None, stats_out=None):
        """R2 in one launch (index draw + assembly); stats_out: optional float32[blocks, 8] buffer for the emotion sums the fast
        update path consumes (returned as the second value)."""
        if out is None:
            out = torch.empty(batch, L.REC, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_replay_sample_gather(C.c_int(batch), C.c_int64(self.size), C.c_uint64(seed),
                                                         C.c_uint32(update_idx & 0xFFFFFFFF), C.c_uint32(stream_id), L.ptr(self.buf),
                                                         L.ptr(idx_out), L.ptr(out), L.ptr(stats_out), L.stream_ptr()),
                    "erlfill_replay_sample_gather")
        return out, stats_out

    def gather(self, idx, out=None):
        b = idx.shape[0]
        if out is None:
            out = torch.empty(b, L.REC, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_replay_gather(C.c_int(b), L.ptr(self.buf), L.ptr(idx), L.ptr(out), L.stream_ptr()),
                    "erlfill_replay_gather")
        return out


class DDPGLearner:
    """Flat-buffer ER-DDPG networks + optimiser state and the fused update (DDPGAgent.learn)."""

    def __init__(self, actor_sd, critic_sd, device="cuda", batch_size=128, gamma=0.99, tau=0.01, actor_lr=1e-4,
                 critic_lr=1e-3, lr_decay_rate=0.999, precision="exact", dropout="off", seed=0, rank=0):
        self.device = torch.device(device)
        d = self.device
        self.actor = FlatNet(ACTOR_LAYOUT, L.ACTOR_PARAMS, d)
        self.actor_target = FlatNet(ACTOR_LAYOUT, L.ACTOR_PARAMS, d)
        self.critic = FlatNet(CRITIC_LAYOUT, L.CRITIC_PARAMS, d)
        self.critic_target = FlatNet(CRITIC_LAYOUT, L.CRITIC_PARAMS, d)
        self.actor.load(actor_sd); self.actor_target.load(actor_sd)
        self.critic.load(critic_sd); self.critic_target.load(critic_sd)
        self.scale = torch.as_tensor(actor_sd["scale_params"]).to(d, torch.float32).contiguous()
        self.offset = torch.as_tensor(actor_sd["offset_params"]).to(d, torch.float32).contiguous()
        self.actor_m = torch.zeros(L.ACTOR_PARAMS, device=d); self.actor_v = torch.zeros(L.ACTOR_PARAMS, device=d)
        self.critic_m = torch.zeros(L.CRITIC_PARAMS, device=d); self.critic_v = torch.zeros(L.CRITIC_PARAMS, device=d)
        self.gamma, self.tau, self.actor_lr, self.critic_lr, self.lr_decay_rate = gamma, tau, actor_lr, critic_lr, lr_decay_rate
        self.n_cond, self.scale_tab, self.offset_tab = 1, None, None      # set_condition_scaling: mixed-condition batches
        self.adam_step = 0
        self.report = torch.zeros(4, dtype=torch.float32, device=d)
        self._ws, self._ws_batch = None, 0
        self._ws_shared = False          # True while a parallel.PeerExchange owns the workspace (peers read its buckets in place)
        if precision not in ("exact", "fast"):
            raise L.ErlfillError("DDPGLearner: precision must be 'exact' (3xTF32, 1e-4 parity) or 'fast' (tcgen05 bf16 operands)")
        self.precision = precision
        # dropout of the four networks inside learn(): "off" = module.eval(); "philox" = train() mode (what the reference runs:
        # it never calls .eval() on this path), keep-masks from the counter-based generator; update(masks=...) injects them
        if dropout not in ("off", "philox"):
            raise L.ErlfillError("DDPGLearner: dropout must be 'off' or 'philox'")
        self.dropout = dropout
        self.seed, self.rank = int(seed), int(rank)
        self.updates = 0                 # update index: Philox counter word of the dropout masks
        self._images_ok = False          # bf16 weight images of the fast path match the fp32 parameters
        self.batch_size = batch_size
        self.use_graph = True
        self._graphs = {}
        self._warm = set()

    def set_condition_scaling(self, scale_tab, offset_tab):
        """Per-condition scale / offset rows [n_cond, 10] for a batch that mixes operating conditions: the update maps the actor
        output of a sampled record with the row of the record's own condition (column 19, written by ReplayBuffer.insert(cond=...))."""
        self.scale_tab = torch.as_tensor(scale_tab).to(self.device, torch.float32).contiguous()
        self.offset_tab = torch.as_tensor(offset_tab).to(self.device, torch.float32).contiguous()
        self.n_cond = int(self.scale_tab.shape[0])
        self._graphs.clear(); self._warm.clear()

    def _workspace(self, batch):
        if self._ws is None or self._ws_batch != batch:
            if self._ws_shared:
                raise L.ErlfillError("DDPGLearner: the workspace is shared with the other ranks (parallel.PeerExchange was built for "
                                     "minibatch %d); a minibatch of %d needs a new PeerExchange" % (self._ws_batch, batch))
            nbytes = L.lib().erlfill_ddpg_workspace_bytes(C.c_int(batch))
            self._ws = torch.zeros((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            self._ws_batch = batch
            # captured graphs hold the old workspace's address (and the old allocation may be reused by anything)
            self._graphs.clear()
            self._warm.clear()
            self._images_ok = False
        return self._ws

    def precision_name(self):
        return {"exact": "f32 (3xTF32 tensor-core GEMMs, fp32 accumulate)",
                "fast": "bf16 operands / f32 accumulate (tcgen05), f32 master weights and optimiser"}[self.precision]

    def launches_per_update(self, data_parallel=False):
        """Kernels of liberlfill.so per erlfill_ddpg_update(PHASE_ALL) (+ the two peer exchanges when data-parallel)."""
        n = int(L.lib().erlfill_ddpg_update_launches(C.c_int(1 if self.precision == "fast" else 0)))
        return n + (int(L.lib().erlfill_ddpg_peer_launches(C.c_int(1 if self.precision == "fast" else 0))) if data_parallel else 0)

    def grad_buckets(self, batch):
        """(critic_bucket, actor_bucket) tensors aliasing the gradient buckets in the workspace."""
        ws = self._workspace(batch)
        cg, ag = C.c_void_p(), C.c_void_p()
        cl, al = C.c_int(), C.c_int()
        L.check(L.lib().erlfill_ddpg_grad_buckets(L.ptr(ws), C.c_int(batch), C.byref(cg), C.byref(cl), C.byref(ag), C.byref(al)),
                "erlfill_ddpg_grad_buckets")
        co = (cg.value - ws.data_ptr()) // 4
        ao = (ag.value - ws.data_ptr()) // 4
        return ws[co:co + cl.value], ws[ao:ao + al.value]

    def exchange_buffers(self, batch):
        """Views of the workspace's exchange buffers: gradient buckets, peer-reduced buckets, [2][64] sum-of-squares partials."""
        ws = self._workspace(batch)
        p = [C.c_void_p() for _ in range(5)]
        cl, al = C.c_int(), C.c_int()
        L.check(L.lib().erlfill_ddpg_exchange_buffers(L.ptr(ws), C.c_int(batch), C.byref(p[0]), C.byref(cl), C.byref(p[1]), C.byref(al),
                                                      C.byref(p[2]), C.byref(p[3]), C.byref(p[4])), "erlfill_ddpg_exchange_buffers")
        o = [(x.value - ws.data_ptr()) // 4 for x in p]
        return {"critic_grad": ws[o[0]:o[0] + cl.value], "actor_grad": ws[o[1]:o[1] + al.value], "critic_reduced": ws[o[2]:o[2] + cl.value],
                "actor_reduced": ws[o[3]:o[3] + al.value], "sumsq": ws[o[4]:o[4] + 128]}

    def actor_struct(self, target=False):
        """erlfill_actor_weights pointing into the flat buffer (rollout kernel reads the live parameters)."""
        v = (self.actor_target if target else self.actor).views
        return L.ActorWeights(L.ptr(v["input_layer.weight"]), L.ptr(v["input_layer.bias"]), L.ptr(v["hidden.2.weight"]),
                              L.ptr(v["hidden.2.bias"]), L.ptr(v["hidden.5.weight"]), L.ptr(v["hidden.5.bias"]),
                              L.ptr(v["hidden.7.weight"]), L.ptr(v["hidden.7.bias"]), L.ptr(v["emotion_weights"]),
                              L.ptr(self.scale), L.ptr(self.offset))

    def params_changed(self):
        """Call after writing the flat parameter buffers from outside (checkpoint load, load_weights): the fast path's bf16
        weight images are rebuilt before the next update."""
        self._images_ok = False

    def _pack_images(self, nets, ws, B):
        with torch.cuda.device(self.device):
            L.check(L.lib().erlfill_ddpg_pack_images(C.byref(nets), L.ptr(ws), C.c_int(B), L.stream_ptr()), "erlfill_ddpg_pack_images")
        self._images_ok = True

    def update(self, batch, next_emotion, train_step, phases=L.PHASE_ALL, allreduce=None, graph=None, masks=None, stats_partials=None):
        """One learn() on a gathered batch [B,20].  allreduce(tensor): optional in-place averaging callback
        applied to the critic bucket and then the actor bucket (data-parallel).

        graph: replay the whole update (about 100 small launches, host-launch bound otherwise) as ONE captured CUDA
        graph.  Default: on for the single-process full update, off with an all-reduce callback or partial phases.
        The graph is keyed on the (batch, next_emotion) buffers; per-step scalars live in the workspace
        (erlfill_ddpg_set_step), so the same graph serves every step."""
        B = batch.shape[0]
        ws = self._workspace(B)
        if phases & L.PHASE_CRITIC_GRAD:
            self.adam_step += 1
        if graph is None:
            graph = self.use_graph and phases == L.PHASE_ALL and (allreduce is None or getattr(allreduce, "graph_safe", False))
        multi = self.n_cond > 1
        nets = L.DDPGNets(L.ptr(self.actor.flat), L.ptr(self.actor_target.flat), L.ptr(self.critic.flat),
                          L.ptr(self.critic_target.flat), L.ptr(self.actor_m), L.ptr(self.actor_v), L.ptr(self.critic_m),
                          L.ptr(self.critic_v), L.ptr(self.scale_tab if multi else self.scale),
                          L.ptr(self.offset_tab if multi else self.offset))
        decay = self.lr_decay_rate ** train_step
        fast = self.precision == "fast"
        if fast and not self._images_ok:
            self._pack_images(nets, ws, B)
        mstruct = None
        if masks is not None:            # ten uint8 keep arrays (include/erlfill.h: erlfill_ddpg_masks), parity runs only
            mstruct = L.DDPGMasks()
            self._mask_refs = [m.contiguous() for m in masks]
            for i, m in enumerate(self._mask_refs):
                mstruct.keep[i] = m.data_ptr()
            graph = False
        dmode = L.DROPOUT_MASK if masks is not None else (L.DROPOUT_PHILOX if self.dropout == "philox" else L.DROPOUT_OFF)
        if dmode == L.DROPOUT_PHILOX:
            graph = False                # the update index is a by-value launch parameter
        if phases & L.PHASE_CRITIC_GRAD:
            self.updates += 1
        upd_idx = (self.updates - 1) & 0xFFFFFFFF      # one index per update, whichever way its phases are called

        # data-parallel over peer memory (parallel.PeerExchange attaches itself to the learner): the fast path's X_GRAD phases raise this
        # rank's "bucket ready" flags themselves, whichever way the phases are called
        px = allreduce if getattr(allreduce, "peer", False) else getattr(self, "_peer_exchange", None)

        def run(ph, adam_step):
            hyper = L.DDPGHyper(B, adam_step, self.gamma, self.tau, 5.0, self.n_cond, self.critic_lr, self.actor_lr, decay,
                                L.PRECISION_FAST if fast else L.PRECISION_EXACT, dmode, self.seed, upd_idx, self.rank,
                                C.cast(C.pointer(mstruct), C.c_void_p) if mstruct is not None else None,
                                L.ptr(stats_partials), 0 if stats_partials is None else int(stats_partials.shape[0]), 0,
                                C.cast(C.pointer(px.g), C.c_void_p) if (fast and px is not None) else None)
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_ddpg_update(C.byref(nets), C.byref(hyper), L.ptr(batch), L.ptr(next_emotion), L.ptr(ws),
                                                    C.c_int(ph), L.ptr(self.report), L.stream_ptr()), "erlfill_ddpg_update")

        # everything a captured graph bakes in: buffer addresses (the workspace included) and the by-value hyper-parameters
        key = (B, batch.data_ptr(), next_emotion.data_ptr(), ws.data_ptr(), id(allreduce), self.gamma, self.tau, self.critic_lr,
               self.actor_lr, self.n_cond, self.precision, 0 if stats_partials is None else stats_partials.data_ptr())
        if graph and key not in self._graphs and key not in self._warm:
            self._warm.add(key)          # first update on these buffers runs eagerly (module loading outside a capture)
            graph = False
        if graph:
            with torch.cuda.device(self.device):
                L.check(L.lib().erlfill_ddpg_set_step(L.ptr(ws), C.c_int(B), C.c_int(self.adam_step), C.c_double(decay),
                                                      L.stream_ptr()), "erlfill_ddpg_set_step")
            g = self._graphs.get(key)
            if g is None:
                if len(self._graphs) >= 4:
                    self._graphs.clear()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    if allreduce is None:
                        run(L.PHASE_ALL, 0)
                    else:       # data-parallel: the two bucket exchanges are captured between the phases
                        self._run_dp(run, allreduce, B, 0)
                self._graphs[key] = g
            g.replay()
        elif allreduce is None:
            run(phases, self.adam_step)
        else:
            self._run_dp(run, allreduce, B, self.adam_step)
        return self.report

    def _run_dp(self, run, allreduce, B, adam_step):
        """GRAD | average the bucket over the ranks | APPLY, for the critic and then the actor."""
        if getattr(allreduce, "peer", False):      # parallel.PeerExchange: one-shot NVLink exchange fused with the norm partials
            red = L.PHASE_REDUCED
            run(L.PHASE_CRITIC_GRAD, adam_step); allreduce.exchange(0); run(L.PHASE_CRITIC_APPLY | red, 0)
            run(L.PHASE_ACTOR_GRAD, 0); allreduce.exchange(1); run(L.PHASE_ACTOR_APPLY | red, 0)
        else:                                      # any in-place averaging callback (parallel.GradSync: NCCL / gloo)
            cb, ab = self.grad_buckets(B)
            run(L.PHASE_CRITIC_GRAD, adam_step); allreduce(cb); run(L.PHASE_CRITIC_APPLY, 0)
            run(L.PHASE_ACTOR_GRAD, 0); allreduce(ab); run(L.PHASE_ACTOR_APPLY, 0)
