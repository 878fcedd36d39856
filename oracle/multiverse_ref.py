# coding=utf-8
"""CPU oracle for the Multiverse encoder-decoder ConvRNN hot path.

THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it.  The product path (``multiverse_b200``) never routes through it.

PINNED ON AN EXECUTION OF THE REFERENCE'S OWN GRAPH CODE (round 2).  The reference
(JunweiLiang/Multiverse @ c1756f0) ships no unit tests, golden vectors or fixtures for this path, and
its arithmetic lives in TensorFlow 1.15 (``tf.contrib.rnn.ConvLSTMCell``, ``tf.nn.raw_rnn``,
``tf.nn.dynamic_rnn``, ``tf.nn.conv2d`` ...), which is neither vendored under the reference tree nor
installable in this image (Python 3.12, no wheel, no network).  So instead of TensorFlow,
``oracle/tf1_eager`` provides an eager torch-fp64 stand-in for the ~90 ``tf.*`` symbols
``code/pred_models.py`` touches, and ``oracle/tf1_eager/run_reference.py`` imports the UNMODIFIED
``/root/reference/code/pred_models.py`` against it and runs ``Model.__init__ / build_forward /
build_loss`` and ``Trainer.__init__`` line by line: scopes and variable names, both ``raw_rnn`` loop
functions, the beam bookkeeping and back-trace, ``add_div_penalty`` / ``gather_helper``, the graph
attention, the loss and ``tf.gradients`` -> clip -> Adadelta are executed reference code.
``tests/test_reference_exec_cpu.py`` asserts that this execution equals this file to 1e-12 with
identical beam ids (greedy two-scale, K=5 plain and K=20 diverse beams, use_gnn off, one training
step incl. every clipped gradient and the Adadelta update), and the committed rollout goldens
(``tests/golden/rollout_*.npz``, ``source = "reference_exec"``) are outputs of that execution.
What stays restated - and is what this file's elementary ops and the stand-in's ops both follow -
are the per-op TensorFlow semantics (SAME padding, ConvLSTMCell gate order and forget bias,
raw_rnn / dynamic_rnn protocol, top_k / argmax tie-breaking, l2_normalize, Huber, Adadelta; sheet in
SURVEY.md section 8c).  Those are anchored outside this repo where PyTorch implements the same op:
torch conv2d with TF's SAME rule, torch.nn.LSTMCell (the cell on a 1x1 grid), torch.optim.Adadelta,
torch's Huber / cross-entropy (tests/test_oracle_cpu.py); plus the literal dense [HW,HW] graph
attention, a brute-force beam replay and an independently written torch port of the whole model.
The evaluation metrics (minADE / minFDE / NLL, row f-3) are pinned directly on the reference's own
numpy functions (tests/golden/make_golden_metrics.py imports them).

Every function cites the reference file:line it follows (paths relative to the
reference root).  All functions are dtype-generic: pass float64 arrays for the
"truth" run and float32 for what TF-CPU would compute (modulo summation order).
Layout is NHWC throughout, exactly like the reference.
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import numpy as np

# --------------------------------------------------------------------------- #
# elementary ops
# --------------------------------------------------------------------------- #


def sigmoid(x):
  """Numerically stable logistic, same dtype as x."""
  out = np.empty_like(x)
  pos = x >= 0
  out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
  ex = np.exp(x[~pos])
  out[~pos] = ex / (1.0 + ex)
  return out


def same_pad(in_size, k, stride):
  """TF "SAME" padding rule (tf.nn.conv2d): out=ceil(in/s),
  pad_total=max((out-1)*s+k-in,0), pad_before=pad_total//2, rest after.
  Stride-2/k=3 on even sizes therefore pads 0 before / 1 after."""
  out = -(-in_size // stride)
  total = max((out - 1) * stride + k - in_size, 0)
  before = total // 2
  return out, before, total - before


def conv2d_same(x, w, stride=1):
  """tf.nn.conv2d(x, w, [1,s,s,1], "SAME"), NHWC / HWIO, cross-correlation.

  Called by the reference helper ``conv2d`` (code/pred_models.py:1333-1373) and
  by ConvLSTMCell's ``_conv`` (TF 1.15 contrib/rnn/python/ops/rnn_cell.py).
  Implemented as im2col + one matmul in x.dtype.
  """
  n, h, wd, c = x.shape
  kh, kw, ci, co = w.shape
  assert ci == c, (x.shape, w.shape)
  oh, pt, pb = same_pad(h, kh, stride)
  ow, pl, pr = same_pad(wd, kw, stride)
  xp = np.zeros((n, h + pt + pb, wd + pl + pr, c), dtype=x.dtype)
  xp[:, pt:pt + h, pl:pl + wd, :] = x
  cols = np.empty((n, oh, ow, kh * kw * c), dtype=x.dtype)
  for dy in range(kh):
    for dx in range(kw):
      patch = xp[:, dy:dy + (oh - 1) * stride + 1:stride,
                 dx:dx + (ow - 1) * stride + 1:stride, :]
      cols[..., (dy * kw + dx) * c:(dy * kw + dx + 1) * c] = patch
  out = cols.reshape(-1, kh * kw * c) @ w.reshape(kh * kw * c, co).astype(x.dtype)
  return out.reshape(n, oh, ow, co)


def conv2d_layer(x, W, b=None, stride=1, activation=None):
  """The reference's ``conv2d`` helper, code/pred_models.py:1333-1373:
  conv SAME, optional bias_add, then activation (tanh in every published
  config, identity for hidden2grid)."""
  y = conv2d_same(x, W, stride)
  if b is not None:
    y = y + b.astype(x.dtype)
  if activation == "tanh":
    y = np.tanh(y)
  elif activation == "relu":
    y = np.maximum(y, 0)
  elif activation == "lrelu":
    y = np.where(y > 0, y, 0.2 * y)
  elif activation is not None:
    raise ValueError(activation)
  return y


def one_hot(ids, depth, dtype):
  """tf.one_hot -> float (code/pred_models.py:174,414,604)."""
  out = np.zeros(ids.shape + (depth,), dtype=dtype)
  np.put_along_axis(out, ids[..., None].astype(np.int64), 1.0, axis=-1)
  return out


# --------------------------------------------------------------------------- #
# a1: ConvLSTM cell
# --------------------------------------------------------------------------- #


def convlstm_cell(x, c, h, kernel, biases, forget_bias=1.0):
  """tf.contrib.rnn.ConvLSTMCell.call (TF 1.15), as built at
  code/pred_models.py:189-202 (encoders) and :236-249 (decoders).

  ``_conv([inputs, hidden])`` concatenates inputs FIRST along channels; the
  variables are ``kernel [kh,kw,Cx+Ch,4Ch]`` and ``biases [4Ch]``; the result
  is split as input_gate, new_input, forget_gate, output_gate;
  ``new_cell = sigmoid(f + forget_bias)*cell + sigmoid(i)*tanh(j)``;
  ``output = tanh(new_cell)*sigmoid(o)``.  Returns (c', h').
  """
  xh = np.concatenate([x, h], axis=-1)
  g = conv2d_same(xh, kernel, 1) + biases.astype(x.dtype)
  gi, gj, gf, go = np.split(g, 4, axis=-1)
  new_c = sigmoid(gf + forget_bias) * c + sigmoid(gi) * np.tanh(gj)
  new_h = np.tanh(new_c) * sigmoid(go)
  return new_c, new_h


# --------------------------------------------------------------------------- #
# a4: scene CNN
# --------------------------------------------------------------------------- #


def scene_cnn(scene_feat, obs_scene, weights, num_scales=2):
  """code/pred_models.py:146-165: embedding_lookup of the per-frame one-hot
  segmentation, then ``tanh(conv3x3 stride 2 SAME + b)`` once per scale.
  Returns a list of [N,T,SH/2^(i+1),SW/2^(i+1),Cs]."""
  n, t = obs_scene.shape
  x = scene_feat[obs_scene.reshape(-1)]  # [N*T,SH,SW,SC]
  outs = []
  for i in range(num_scales):
    x = conv2d_layer(x, weights["person_pred/scene_conv%d/W" % (i + 1)],
                     weights["person_pred/scene_conv%d/b" % (i + 1)],
                     stride=2, activation="tanh")
    outs.append(x.reshape((n, t) + x.shape[1:]))
  return outs


# --------------------------------------------------------------------------- #
# a7-a9: graph attention
# --------------------------------------------------------------------------- #


def l2_normalize(x, eps=1e-12):
  """tf.nn.l2_normalize(x, -1): x * rsqrt(max(sum(x^2), eps))."""
  ss = np.sum(x * x, axis=-1, keepdims=True)
  return x / np.sqrt(np.maximum(ss, eps))


def neighbour_mask(h, w, dtype):
  """code/pred_models.py:885-902: conv2d(one_hot(HW) as [HW,H,W,1], ones[3,3],
  SAME) -> [H,W,H,W] with 1 at self + 8 neighbours (clipped at borders)."""
  eye = np.eye(h * w, dtype=dtype).reshape(h * w, h, w, 1)
  m = conv2d_same(eye, np.ones((3, 3, 1, 1), dtype=dtype), 1)
  return m.reshape(h, w, h, w)


def gnn_dense(hstate, scene_mean):
  """LITERAL restatement of gnn_edge (code/pred_models.py:808-858),
  gnn_mask_edge with exp_mask (:885-909, :1399-1401) and gnn_node (:860-882),
  plus the residual (:378, :651): the full [HW,HW] cosine matrix, -1e30 on
  non-neighbours, softmax over HW, matmul with the states.  O((HW)^2)."""
  n, h, w, d = hstate.shape
  dt = hstate.dtype
  feats = hstate.reshape(n, h * w, d)
  if scene_mean is not None:
    feats = np.concatenate([feats, scene_mean.reshape(n, h * w, -1)], axis=-1)
  fn = l2_normalize(feats)
  edge = fn @ np.transpose(fn, (0, 2, 1))  # [N,HW,HW]
  mask = neighbour_mask(h, w, dt).reshape(h * w, h * w)
  edge = edge + (1.0 - mask)[None] * dt.type(-1e30)
  edge = edge - edge.max(axis=-1, keepdims=True)
  e = np.exp(edge)
  a = e / e.sum(axis=-1, keepdims=True)
  summed = a @ hstate.reshape(n, h * w, d)
  return hstate + summed.reshape(n, h, w, d)


def gnn_stencil(hstate, scene_mean):
  """Same result as ``gnn_dense`` computed on the 3x3 band only (what the CUDA
  kernel does).  tests/ prove the two agree."""
  n, h, w, d = hstate.shape
  dt = hstate.dtype
  feats = hstate
  if scene_mean is not None:
    feats = np.concatenate([hstate, scene_mean], axis=-1)
  fn = l2_normalize(feats)
  fp = np.zeros((n, h + 2, w + 2, fn.shape[-1]), dtype=dt)
  fp[:, 1:-1, 1:-1] = fn
  hp = np.zeros((n, h + 2, w + 2, d), dtype=dt)
  hp[:, 1:-1, 1:-1] = hstate
  valid = np.zeros((h + 2, w + 2), dtype=bool)
  valid[1:-1, 1:-1] = True
  scores = np.full((n, h, w, 9), -np.inf, dtype=dt)
  for k, (dy, dx) in enumerate([(a, b) for a in range(3) for b in range(3)]):
    s = np.sum(fn * fp[:, dy:dy + h, dx:dx + w], axis=-1)
    ok = valid[dy:dy + h, dx:dx + w]
    scores[..., k] = np.where(ok[None], s, -np.inf)
  scores = scores - scores.max(axis=-1, keepdims=True)
  e = np.exp(scores)
  a = e / e.sum(axis=-1, keepdims=True)
  out = hstate.copy()
  for k, (dy, dx) in enumerate([(a_, b_) for a_ in range(3) for b_ in range(3)]):
    out += a[..., k:k + 1] * hp[:, dy:dy + h, dx:dx + w]
  return out


# --------------------------------------------------------------------------- #
# a10/a11: grid embedding and heads
# --------------------------------------------------------------------------- #


def grid_emb(x, W, b):
  """Model.grid_emb (code/pred_models.py:912-919): tanh(conv3x3(x)+b)."""
  return conv2d_layer(x, W, b, stride=1, activation="tanh")


def hidden2grid(hstate, W):
  """Model.hidden2grid (code/pred_models.py:925-959): conv3x3, no bias."""
  return conv2d_layer(hstate, W, None, stride=1, activation=None)


# --------------------------------------------------------------------------- #
# a3: encoders (tf.nn.dynamic_rnn, zero initial state, no masking since
#     sequence_length == T for every row, code/pred_models.py:1057-1063)
# --------------------------------------------------------------------------- #


def encoder(inputs, kernel, biases, ch):
  """code/pred_models.py:212-215 / :232-234.  inputs [N,T,H,W,Cx]."""
  n, t, h, w, _ = inputs.shape
  c = np.zeros((n, h, w, ch), dtype=inputs.dtype)
  hs = np.zeros((n, h, w, ch), dtype=inputs.dtype)
  for step in range(t):
    c, hs = convlstm_cell(inputs[:, step], c, hs, kernel, biases)
  return c, hs


# --------------------------------------------------------------------------- #
# a5: greedy decoder
# --------------------------------------------------------------------------- #


def grid_decoder(first_input, enc_state, pred_len, cell_w, emb_w, head_w,
                 scene_mean=None, use_gnn=False, input_onehot=False,
                 return_steps=False):
  """Model.grid_decoder (code/pred_models.py:311-471) under tf.nn.raw_rnn
  (:455) at inference (no teacher forcing), SURVEY.md appendix A.1.

  cell_w=(kernel,biases), emb_w=(W,b), head_w=W.
  Returns (decoder_out [N,Tp,H,W,P], decoder_out_h [N,Tp,H,W,Ch]) and, when
  return_steps, the per-step (c,h) list too.
  """
  c, h = enc_state
  n, hh, ww, p = first_input.shape
  inp = first_input
  out_h = []
  steps = []
  for _ in range(pred_len):
    h_in = gnn_dense(h, scene_mean) if use_gnn else h     # :359-382
    x = grid_emb(inp, emb_w[0], emb_w[1])                 # :442-446
    c, h = convlstm_cell(x, c, h_in, cell_w[0], cell_w[1])
    out_h.append(h)
    steps.append((c, h))
    o = hidden2grid(h, head_w)                            # :422-425 / :432-435
    if input_onehot:
      ids = np.argmax(o.reshape(n, hh * ww), axis=1)      # :411-415
      inp = one_hot(ids, hh * ww, h.dtype).reshape(n, hh, ww, 1)
    else:
      inp = o
  out_h = np.stack(out_h, axis=1)
  dec_out = hidden2grid(out_h.reshape((-1,) + out_h.shape[2:]), head_w)  # :467
  dec_out = dec_out.reshape(n, pred_len, hh, ww, p)
  if return_steps:
    return dec_out, out_h, steps
  return dec_out, out_h


# --------------------------------------------------------------------------- #
# a6: beam decoder
# --------------------------------------------------------------------------- #


def log_softmax(x):
  m = x.max(axis=-1, keepdims=True)
  s = x - m
  return s - np.log(np.sum(np.exp(s), axis=-1, keepdims=True))


def rank_within_row(lp):
  """add_div_penalty steps 1-2 (code/pred_models.py:1210-1217):
  top_k(k=V, sorted) then invert_permutation -> rank of every entry, 0 = largest,
  ties broken towards the lower index (tf.nn.top_k is stable)."""
  order = np.argsort(-lp, axis=-1, kind="stable")
  rank = np.empty_like(order)
  np.put_along_axis(rank, order, np.arange(lp.shape[-1])[None, None, :]
                    * np.ones_like(order), axis=-1)
  return rank


def add_div_penalty(lp, gamma):
  """code/pred_models.py:1197-1223."""
  return lp + lp.dtype.type(math.log(gamma)) * rank_within_row(lp).astype(lp.dtype)


def top_k_sorted(x, k):
  """tf.nn.top_k(x, k, sorted=True): descending, ties -> lower index first."""
  order = np.argsort(-x, axis=-1, kind="stable")[..., :k]
  return np.take_along_axis(x, order, axis=-1), order.astype(np.int32)


def grid_decoder_beam_search(first_input, enc_state, pred_len, beam_size,
                             cell_w, emb_w, head_w, scene_mean=None,
                             use_gnn=False, diverse_beam=False, diverse_gamma=1.0,
                             fix_num_timestep=0, return_trace=False):
  """Model.grid_decoder_beam_search (code/pred_models.py:474-806), SURVEY.md
  appendix A.2.  Returns (best_beam_logits [N,Tp,H,W,1], logits [N,B,Tp,V],
  ids [N,B,Tp] int32, logprobs [N,B]) and, when return_trace, the raw per-step
  (ids, parents, logits, scores) before back-trace."""
  c0, h0 = enc_state
  n, hh, ww, ch = h0.shape
  v = hh * ww
  b = beam_size
  dt = h0.dtype
  rep = lambda t: np.repeat(t[:, None], b, axis=1).reshape((n * b,) + t.shape[1:])
  c, h = rep(c0), rep(h0)                                   # :499-502,:527-529
  inp = rep(first_input)                                    # :497,:531
  sm = rep(scene_mean) if scene_mean is not None else None  # :831-834
  score = np.zeros((n, b), dtype=dt)                        # :514
  step_ids, step_parents, step_logits, step_scores = [], [], [], []

  def cell_step(inp, c, h):
    h_in = gnn_dense(h, sm) if use_gnn else h               # :631-654
    x = grid_emb(inp, emb_w[0], emb_w[1])                   # :662-666
    return convlstm_cell(x, c, h_in, cell_w[0], cell_w[1])

  c, h = cell_step(inp, c, h)                               # loop_fn(time=0) + cell
  for time in range(1, pred_len + 1):
    logits = hidden2grid(h, head_w).reshape(n, b, v)        # :550-555
    lp = log_softmax(logits) + score[:, :, None]            # :557-560
    if diverse_beam:
      lp = add_div_penalty(lp, diverse_gamma)               # :561-567
    cand = lp.reshape(n, b * v) if time > 1 else lp[:, 0]   # :569-573
    new_score, idx = top_k_sorted(cand, b)                  # :578
    if time <= fix_num_timestep:                            # :581-584
      new_score = np.zeros_like(new_score)
    ids = (idx % v).astype(np.int32)                        # :588
    parents = (idx // v).astype(np.int32)                   # :591
    step_ids.append(ids)
    step_parents.append(parents)
    step_logits.append(logits)
    step_scores.append(new_score)
    score = new_score
    flat = (parents + (np.arange(n) * b)[:, None]).reshape(-1)  # gather_helper :1235
    c, h = c[flat], h[flat]                                 # :611-623
    inp = one_hot(ids.reshape(-1), v, dt).reshape(n * b, hh, ww, 1)  # :602-606
    if time == pred_len:
      break
    c, h = cell_step(inp, c, h)
  # back-trace, :689-764
  par = np.tile(np.arange(b, dtype=np.int32)[None], (n, 1))  # :714-716
  out_ids = np.zeros((n, b, pred_len), dtype=np.int32)
  out_logits = np.zeros((n, b, pred_len, v), dtype=dt)
  rows = np.arange(n)[:, None]
  for tau in range(pred_len - 1, -1, -1):
    out_ids[:, :, tau] = step_ids[tau][rows, par]
    out_logits[:, :, tau] = step_logits[tau][rows, par]
    par = step_parents[tau][rows, par]
  best = out_logits[:, 0].reshape(n, pred_len, hh, ww, 1)    # :799-803
  if return_trace:
    return best, out_logits, out_ids, score, dict(
        ids=step_ids, parents=step_parents, logits=step_logits, scores=step_scores)
  return best, out_logits, out_ids, score


# --------------------------------------------------------------------------- #
# a12: loss
# --------------------------------------------------------------------------- #


def huber(pred, target, delta=1.0):
  """tf.losses.huber_loss(reduction=MEAN), code/pred_models.py:1016-1022."""
  e = np.abs(pred - target)
  q = np.minimum(e, delta)
  return np.mean(0.5 * q * q + delta * (e - q))


def sparse_ce(logits, labels):
  """mean(sparse_softmax_cross_entropy_with_logits), :991-995."""
  lp = log_softmax(logits)
  return -np.mean(lp[np.arange(logits.shape[0]), labels])


# --------------------------------------------------------------------------- #
# whole forward (Model.build_forward, code/pred_models.py:123-308)
# --------------------------------------------------------------------------- #


def default_config(**kw):
  """The canonical hyper-parameters of every published command
  (TRAINING.md:32-39, TESTING.md:32-39,84-93) at the BASELINE.json grid shape."""
  cfg = dict(
      batch_size=4, obs_len=8, pred_len=12, scene_h=72, scene_w=36,
      scene_class=11, scene_conv_dim=64, scene_conv_kernel=3,
      scene_grid_strides=[2, 4], use_grids=[True, True],
      enc_hidden_size=256, dec_hidden_size=256, emb_size=32,
      convlstm_kernel=3, use_scene_enc=True, use_gnn=True,
      use_beam_search=False, beam_size=1, diverse_beam=False,
      diverse_gamma=1.0, fix_num_timestep=0, activation_func="tanh",
      video_h=1080, video_w=1920)
  cfg.update(kw)
  cfg = SimpleNamespace(**cfg)
  # code/pred_utils.py:127-132
  cfg.scene_grids = [(int(round(cfg.scene_h * 1.0 / s)), int(round(cfg.scene_w * 1.0 / s)))
                     for s in cfg.scene_grid_strides]
  return cfg


def weight_shapes(cfg):
  """Variable names/shapes as the reference creates them (SURVEY.md §8a
  'Weights'), scopes from code/pred_models.py:140,160,193,200,215,234,240,247,
  327,456,446,469,930-950."""
  k = cfg.convlstm_kernel
  ch, e, cs = cfg.enc_hidden_size, cfg.emb_size, cfg.scene_conv_dim
  sk = cfg.scene_conv_kernel
  shp = {}
  cin = cfg.scene_class
  for i in range(len(cfg.scene_grid_strides)):
    shp["person_pred/scene_conv%d/W" % (i + 1)] = (sk, sk, cin, cs)
    shp["person_pred/scene_conv%d/b" % (i + 1)] = (cs,)
    cin = cs
  for i in range(len(cfg.scene_grids)):
    if not cfg.use_grids[i]:
      continue
    cx = cs if cfg.use_scene_enc else e
    p = "person_pred/"
    shp[p + "encoder_grid_class_%d/enc_grid_%d/kernel" % (i, i)] = (k, k, cx + ch, 4 * ch)
    shp[p + "encoder_grid_class_%d/enc_grid_%d/biases" % (i, i)] = (4 * ch,)
    shp[p + "encoder_grid_reg_%d/enc_grid_regress_%d/kernel" % (i, i)] = (k, k, 2 + ch, 4 * ch)
    shp[p + "encoder_grid_reg_%d/enc_grid_regress_%d/biases" % (i, i)] = (4 * ch,)
    for kind, cell, pdim in (("class", "dec_grid_%d" % i, 1), ("reg", "dec_grid_reg_%d" % i, 2)):
      d = p + "decoder_grid_%s_%d/decoder_rnn/" % (kind, i)
      shp[d + cell + "/kernel"] = (k, k, e + ch, 4 * ch)
      shp[d + cell + "/biases"] = (4 * ch,)
      shp[d + "grid_emb/W"] = (3, 3, pdim, e)
      shp[d + "grid_emb/b"] = (e,)
      shp[p + "hidden2grid_decoder_grid_%s_%d/out_dec_grid/W" % (kind, i)] = (3, 3, ch, pdim)
  return shp


def make_weights(cfg, seed=0, dtype=np.float32, head_scale=4.0, bias_scale=0.05):
  """Seeded synthetic weights under the TF variable names:
  ConvLSTM kernels glorot-uniform; conv2d ``W`` ~ N(0, 2/fan_in) clipped at 2
  sigma; ``out_dec_grid/W`` scaled so logit margins are healthy.  Biases get a
  small random value (the reference initialises them to 0 but trained
  checkpoints do not keep them there, and 0 would hide bias-indexing bugs)."""
  rng = np.random.default_rng(seed)
  out = {}
  for name, shp in weight_shapes(cfg).items():
    if name.endswith("/kernel"):
      fan_in = shp[0] * shp[1] * shp[2]
      fan_out = shp[0] * shp[1] * shp[3]
      lim = math.sqrt(6.0 / (fan_in + fan_out))
      w = rng.uniform(-lim, lim, size=shp)
    elif name.endswith("/W"):
      fan_in = shp[0] * shp[1] * shp[2]
      w = np.clip(rng.standard_normal(shp), -2, 2) * math.sqrt(2.0 / fan_in)
      if "out_dec_grid" in name:
        w = w * head_scale
    else:
      w = rng.standard_normal(shp) * bias_scale
    out[name] = w.astype(dtype)
  return out


def grid_centers(cfg):
  """code/multifuture_inference.py:101-113 / code/preprocess.py:99-106."""
  centers = []
  for h, w in cfg.scene_grids:
    h_gap, w_gap = cfg.video_h * 1.0 / h, cfg.video_w * 1.0 / w
    cx = np.cumsum([w_gap] * w) - w_gap / 2.0
    cy = np.cumsum([h_gap] * h) - h_gap / 2.0
    centers.append(np.stack((np.tile(cx[None], (h, 1)), np.tile(cy[:, None], (1, w))), axis=-1))
  return centers


def traj_to_grid(cfg, traj):
  """code/multifuture_inference.py:115-156 (== code/preprocess.py:436-475):
  traj [T,2] pixels -> per scale (class ids [T], offsets [T,h,w,2])."""
  classes, targets = [], []
  for center, (h, w) in zip(grid_centers(cfg), cfg.scene_grids):
    h_gap, w_gap = cfg.video_h * 1.0 / h, cfg.video_w * 1.0 / w
    xi = np.ceil(traj[:, 0] / w_gap).astype(int)
    yi = np.ceil(traj[:, 1] / h_gap).astype(int)
    xi[xi == 0] = 1
    yi[yi == 0] = 1
    xi, yi = xi - 1, yi - 1
    classes.append((yi * w + xi).astype(np.int32))
    targets.append((traj[:, None, None, :] - center[None]).astype(np.float32))
  return classes, targets


def make_inputs(cfg, seed=0, dtype=np.float32):
  """Seeded synthetic feeds of the reference's placeholder shapes
  (code/pred_models.py:62-115): piecewise-constant one-hot scene
  segmentation, one frame per sample, smooth random trajectories in a
  1920x1080 frame, obs targets computed exactly like get_grid_input."""
  rng = np.random.default_rng(seed)
  n, t, tp = cfg.batch_size, cfg.obs_len, cfg.pred_len
  sh, sw, sc = cfg.scene_h, cfg.scene_w, cfg.scene_class
  bh, bw = 6, 3
  blocks = rng.integers(0, sc, size=(n, -(-sh // bh), -(-sw // bw)))
  seg = np.repeat(np.repeat(blocks, bh, axis=1), bw, axis=2)[:, :sh, :sw]
  scene_feat = one_hot(seg, sc, dtype)
  obs_scene = np.tile(np.arange(n, dtype=np.int32)[:, None], (1, t))
  start = rng.uniform([0.2 * cfg.video_w, 0.2 * cfg.video_h],
                      [0.8 * cfg.video_w, 0.8 * cfg.video_h], size=(n, 2))
  vel = rng.normal(0, 25.0, size=(n, t + tp, 2))
  traj = np.clip(start[:, None] + np.cumsum(vel, axis=1), 1.0,
                 [cfg.video_w - 1.0, cfg.video_h - 1.0])
  feeds = dict(scene_feat=scene_feat, obs_scene=obs_scene, traj=traj.astype(dtype),
               grid_obs_labels=[], grid_obs_regress=[], grid_pred_labels=[],
               grid_pred_regress=[])
  ns = len(cfg.scene_grids)
  obs_l = [[] for _ in range(ns)]; obs_r = [[] for _ in range(ns)]
  pr_l = [[] for _ in range(ns)]; pr_r = [[] for _ in range(ns)]
  for i in range(n):
    cl, tg = traj_to_grid(cfg, traj[i])
    for s in range(ns):
      obs_l[s].append(cl[s][:t]); obs_r[s].append(tg[s][:t])
      pr_l[s].append(cl[s][t:]); pr_r[s].append(tg[s][t:])
  for s in range(ns):
    feeds["grid_obs_labels"].append(np.stack(obs_l[s]).astype(np.int32))
    feeds["grid_obs_regress"].append(np.stack(obs_r[s]).astype(dtype))
    feeds["grid_pred_labels"].append(np.stack(pr_l[s]).astype(np.int32))
    feeds["grid_pred_regress"].append(np.stack(pr_r[s]).astype(dtype))
  return feeds


def cast_tree(tree, dtype):
  if isinstance(tree, dict):
    return {k: cast_tree(v, dtype) for k, v in tree.items()}
  if isinstance(tree, (list, tuple)):
    return type(tree)(cast_tree(v, dtype) for v in tree)
  if isinstance(tree, np.ndarray) and tree.dtype.kind == "f":
    return tree.astype(dtype)
  return tree


def scale_weights(weights, i):
  """Pick the per-scale weight tuples out of the TF-named dict."""
  p = "person_pred/"
  g = lambda n: weights[p + n]
  return SimpleNamespace(
      enc_class=(g("encoder_grid_class_%d/enc_grid_%d/kernel" % (i, i)),
                 g("encoder_grid_class_%d/enc_grid_%d/biases" % (i, i))),
      enc_reg=(g("encoder_grid_reg_%d/enc_grid_regress_%d/kernel" % (i, i)),
               g("encoder_grid_reg_%d/enc_grid_regress_%d/biases" % (i, i))),
      dec_class=(g("decoder_grid_class_%d/decoder_rnn/dec_grid_%d/kernel" % (i, i)),
                 g("decoder_grid_class_%d/decoder_rnn/dec_grid_%d/biases" % (i, i))),
      dec_reg=(g("decoder_grid_reg_%d/decoder_rnn/dec_grid_reg_%d/kernel" % (i, i)),
               g("decoder_grid_reg_%d/decoder_rnn/dec_grid_reg_%d/biases" % (i, i))),
      emb_class=(g("decoder_grid_class_%d/decoder_rnn/grid_emb/W" % i),
                 g("decoder_grid_class_%d/decoder_rnn/grid_emb/b" % i)),
      emb_reg=(g("decoder_grid_reg_%d/decoder_rnn/grid_emb/W" % i),
               g("decoder_grid_reg_%d/decoder_rnn/grid_emb/b" % i)),
      head_class=g("hidden2grid_decoder_grid_class_%d/out_dec_grid/W" % i),
      head_reg=g("hidden2grid_decoder_grid_reg_%d/out_dec_grid/W" % i))


def forward(cfg, weights, feeds, dtype=np.float64, return_intermediates=False):
  """Model.build_forward (code/pred_models.py:123-308) at inference:
  scene CNN -> per scale {class encoder, reg encoder, class decoder (greedy or
  beam), reg decoder}.  Returns a dict with ``grid_pred_decoded`` /
  ``grid_pred_reg_decoded`` (lists per scale, [] for unused scales, :170-171)
  and ``beam_outputs`` ([logits, ids, logprobs], :276) or None."""
  weights = cast_tree(weights, dtype)
  feeds = cast_tree(feeds, dtype)
  n = cfg.batch_size
  assert cfg.use_scene_enc, "oracle covers the published use_scene_enc configs"
  scene_convs = scene_cnn(feeds["scene_feat"], feeds["obs_scene"], weights,
                          len(cfg.scene_grid_strides))
  res = dict(grid_pred_decoded=[], grid_pred_reg_decoded=[], beam_outputs=None,
             scene_convs=scene_convs, inter=[])
  for i, (h, w) in enumerate(cfg.scene_grids):
    if not cfg.use_grids[i]:
      res["grid_pred_decoded"].append([])
      res["grid_pred_reg_decoded"].append([])
      res["inter"].append(None)
      continue
    sw = scale_weights(weights, i)
    labels = feeds["grid_obs_labels"][i]
    obs_onehot = one_hot(labels, h * w, dtype).reshape(n, -1, h, w, 1)      # :174-175
    obs_reg = feeds["grid_obs_regress"][i]                                  # :181
    enc_in = scene_convs[i] * obs_onehot                                     # :210
    enc_state = encoder(enc_in, sw.enc_class[0], sw.enc_class[1], cfg.enc_hidden_size)
    enc_reg_state = encoder(obs_reg, sw.enc_reg[0], sw.enc_reg[1], cfg.enc_hidden_size)
    scene_mean = scene_convs[i].mean(axis=1)                                 # :826-828
    if cfg.use_beam_search:
      best, logits, ids, logprobs = grid_decoder_beam_search(
          obs_onehot[:, -1], enc_state, cfg.pred_len, cfg.beam_size,
          sw.dec_class, sw.emb_class, sw.head_class, scene_mean=scene_mean,
          use_gnn=cfg.use_gnn, diverse_beam=cfg.diverse_beam,
          diverse_gamma=cfg.diverse_gamma, fix_num_timestep=cfg.fix_num_timestep)
      res["beam_outputs"] = [logits, ids, logprobs]
      dec, dec_h = best, None
    else:
      dec, dec_h = grid_decoder(obs_onehot[:, -1], enc_state, cfg.pred_len,
                                sw.dec_class, sw.emb_class, sw.head_class,
                                scene_mean=scene_mean, use_gnn=cfg.use_gnn,
                                input_onehot=True)
    reg, reg_h = grid_decoder(obs_reg[:, -1], enc_reg_state, cfg.pred_len,
                              sw.dec_reg, sw.emb_reg, sw.head_reg,
                              use_gnn=False, input_onehot=False)
    res["grid_pred_decoded"].append(dec)
    res["grid_pred_reg_decoded"].append(reg)
    res["inter"].append(dict(enc_state=enc_state, enc_reg_state=enc_reg_state,
                             dec_h=dec_h, reg_h=reg_h, scene_mean=scene_mean)
                        if return_intermediates else None)
  return res


def ids_to_traj(cfg, scale, ids, reg):
  """code/multifuture_inference.py:504-517 / code/pred_utils.py:460-492:
  trajectory point = centre[cell] + offset[cell].  ids [...,Tp], reg
  [Tp,h,w,2] for one sample -> [...,Tp,2]."""
  h, w = cfg.scene_grids[scale]
  centers = grid_centers(cfg)[scale].reshape(h * w, 2)
  reg = reg.reshape(reg.shape[0], h * w, 2)
  t = np.arange(reg.shape[0])
  return centers[ids] + reg[t, ids]
