"""Teacher-forced forward + backward (cxrm_train_step) against torch.autograd on the oracle.

* cross-entropy step, every decoder parameter: gradients vs autograd of the oracle's teacher-forced forward
  (fp32: 2e-3 relative L2 per tensor; bf16 - activations AND activation gradients in bf16, fp32 accumulation - is
  reported and bounded at 0.25: measured worst 0.18 on a self-attention query weight, loss within 5e-4).
* REINFORCE step, LoRA only: the reference backpropagates THROUGH the sampled rollout (scst/gen_prompt.py:279,331-366);
  the oracle does exactly that (autograd through the K/V-cached decode loop), the engine recomputes the sampled
  sequence teacher-forced.  Same loss, same LoRA gradients.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

PAD, BOS, EOS, SEP, PMT_SEP = 4, 1, 2, 3, 9


@pytest.fixture(scope="module")
def sd():
    from oracle import weights
    return weights.make_cxrmate_weights(seed=0)


def _engine(sd, dtype, **kw):
    from cxrmate_b200.engine import Engine
    args = dict(dtype=dtype, image_size=64, max_studies=2, max_images=2, max_prompt=8, max_new_tokens=10, rwd_layers=0,
                enc_chunk=4, max_train_tokens=64)
    args.update(kw)
    e = Engine(**args)
    e.load_state_dict(sd)
    e.finalize()
    return e


def _inputs():
    g = torch.Generator().manual_seed(31)
    px = torch.randn(2, 2, 3, 64, 64, generator=g)
    px[1, 1] = 0.0
    L = 16
    ids = torch.randint(12, 30000, (2, L), generator=g)
    ids[:, 0] = BOS
    ids[0, 7] = SEP
    ids[1, 12:] = PAD                       # right padding inside the teacher-forced batch
    labels = torch.roll(ids, -1, 1)
    labels[:, -1] = EOS
    labels[1, 11:] = PAD
    mask = (ids != PAD).long()
    return px, ids, labels, mask


def _rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-12)).item()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_cross_entropy_backward_all_parameters_vs_autograd(sd, dtype):
    from cxrmate_b200.modelling import position_ids_from_mask, token_ids_to_token_type_ids
    from oracle import bert, cvt
    px, ids, labels, mask = _inputs()
    tt = token_ids_to_token_type_ids(ids, [BOS, SEP], [0, 1, 0])
    pos = position_ids_from_mask(mask)
    # ---- oracle: autograd over the decoder parameters (LoRA factors are part of the forward; gradients of the base weights)
    osd = {k: v.clone() for k, v in sd.items()}
    names = [k for k in osd if k.startswith("decoder.") and "lora_" not in k and k not in
             ("decoder.cls.predictions.decoder.weight", "decoder.cls.predictions.decoder.bias")]
    for k in names:
        osd[k].requires_grad_(True)
    with torch.no_grad():
        mem, mmask = cvt.encode_multi(sd, px)
    logits = bert.decoder_logits(osd, ids, tt, pos, mask, mem, mmask)
    loss_ref = torch.nn.functional.cross_entropy(logits.permute(0, 2, 1), labels, ignore_index=PAD)
    gref = dict(zip(names, torch.autograd.grad(loss_ref, [osd[k] for k in names], allow_unused=True)))
    # ---- engine
    e = _engine(sd, dtype)
    try:
        e.encode(px.cuda())
        e.prefill_cross_kv()
        loss, flat = e.train_step(ids.cuda(), tt.cuda(), pos.cuda(), mask.cuda(), labels.cuda(), loss_kind="ce", ignore_index=PAD,
                                  lora_only=False)
        torch.cuda.synchronize()
        tol_loss, tol = (1e-4, 2e-3) if dtype == "fp32" else (2e-2, 0.25)
        print(f"[{dtype}] CE loss {loss.item():.6f} oracle {loss_ref.item():.6f}")
        assert abs(loss.item() - loss_ref.item()) < tol_loss * max(1.0, abs(loss_ref.item()))
        got = e.grads_by_name(flat, False)
        assert set(got) == set(names), (set(names) ^ set(got))
        worst = ("", 0.0)
        for k in names:
            r = gref[k]
            if r is None:
                r = torch.zeros_like(osd[k])
            g_ = got[k].cpu().reshape(r.shape)
            if r.norm() < 1e-6:      # e.g. key biases: softmax is shift invariant, the true gradient is 0 up to rounding noise
                assert (g_ - r).norm().item() < (1e-5 if dtype == "fp32" else 1e-2), (k, g_.norm().item(), r.norm().item())
                continue
            err = _rel(g_, r)
            if err > worst[1]:
                worst = (k, err)
            assert err < tol, (k, err)
        print(f"[{dtype}] {len(names)} gradient tensors, worst relative L2 error {worst[1]:.2e} ({worst[0]})")
        # staged execution (the all-reduce overlap path) gives the same buffer
        flat2 = torch.zeros_like(flat)
        keep = None
        for st in range(e.train_stages):
            _, flat2 = e.train_step(ids.cuda(), tt.cuda(), pos.cuda(), mask.cuda(), labels.cuda(), loss_kind="ce", ignore_index=PAD,
                                    lora_only=False, grads=flat2, stage=st, _keep=keep)
            keep = e._train_keep
        torch.cuda.synchronize()
        assert _rel(flat2, flat) < 1e-5
    finally:
        e.close()


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_reinforce_backward_lora_vs_autograd_through_the_rollout(sd, dtype):
    """oracle: gradient of reinforce_loss through the sampled K/V-cached rollout (what the reference does);
    engine: grad-free rollout, then one teacher-forced pass with the REINFORCE head"""
    from cxrmate_b200.training import reinforce_backward
    from oracle import cvt, decode, scst
    px, _, _, _ = _inputs()
    prompt = torch.tensor([[8, 500, 9, 600, 601, 1], [8, 10, 9, 11, 1, 4]])
    P, T = prompt.shape[1], 10
    noise = torch.empty(T, 2, 30000).exponential_(1, generator=torch.Generator().manual_seed(9))
    adv = torch.tensor([0.4, -0.7])
    osd = {k: v.clone() for k, v in sd.items()}
    lnames = [k for k in osd if "lora_" in k]
    for k in lnames:
        osd[k].requires_grad_(True)
    with torch.no_grad():
        mem, mmask = cvt.encode_multi(sd, px)
    o = decode.rollout(osd, mem, mmask, prompt, special_token_ids=[BOS, SEP], sections=[0, 1, 0], mask_token_id=PAD,
                       max_new_tokens=T, eos_token_id=EOS, pad_token_id=PAD, do_sample=True, top_k=50, exp_noise=noise)
    loss_ref = scst.reinforce_loss(torch.stack(o.scores, dim=-1), o.sequences[:, P:], adv)
    gref = dict(zip(lnames, torch.autograd.grad(loss_ref, [osd[k] for k in lnames])))
    e = _engine(sd, dtype)
    try:
        e.encode(px.cuda())
        e.prefill_cross_kv()
        if dtype == "fp32":
            out = e.rollout(prompt.cuda(), mode="sample", max_new_tokens=T, eos_token_id=EOS, pad_token_id=PAD, mask_token_id=PAD,
                            special_sample=[BOS, SEP], sections_sample=[0, 1, 0], top_k=50, exp_noise=noise.cuda())
            torch.cuda.synchronize()
            assert torch.equal(out.sequences.cpu(), o.sequences)
            seq = out.sequences
        else:
            seq = o.sequences.cuda()           # bf16 may sample another token at a near-tie: score the oracle's sequence
        loss, flat = reinforce_backward(e, seq, P, adv.cuda(), special_token_ids=[BOS, SEP], sections=[0, 1, 0],
                                        pad_token_id=PAD, top_k=50, lora_only=True)
        torch.cuda.synchronize()
        # bf16: the top-50 survivor set of a flat random-init row differs between bf16 and fp32 logits (a discontinuity of the
        # loss itself), on top of bf16 activation gradients: measured loss 4 %, worst gradient tensor 0.33 on 20 tokens
        tol_loss, tol = (1e-4, 3e-3) if dtype == "fp32" else (8e-2, 0.5)
        print(f"[{dtype}] REINFORCE loss {loss.item():.6f} oracle (through the rollout) {loss_ref.item():.6f}")
        assert abs(loss.item() - loss_ref.item()) < tol_loss * max(1.0, abs(loss_ref.item()))
        got = e.grads_by_name(flat, True)
        assert set(got) == set(lnames)
        worst = 0.0
        for k in lnames:
            err = _rel(got[k].cpu().reshape(gref[k].shape), gref[k])
            worst = max(worst, err)
            assert err < tol, (k, err)
        print(f"[{dtype}] {len(lnames)} LoRA gradient tensors, worst relative L2 error {worst:.2e}")
        if dtype == "fp32":     # the engine's own recorded log-probs give the same loss (cxrm_reinforce_loss)
            l2 = e.reinforce_loss(out.logprobs, adv.cuda())
            assert abs(l2.item() - loss.item()) < 1e-4 * max(1.0, abs(loss.item()))
    finally:
        e.close()


def test_train_step_argument_checks(sd):
    from cxrmate_b200.engine import Engine
    e = Engine(dtype="fp32", image_size=64, max_studies=2, max_images=2, max_prompt=8, max_new_tokens=4, rwd_layers=0)
    try:
        e.load_state_dict(sd)
        e.finalize()
        z = torch.zeros(2, 8, dtype=torch.int64, device="cuda")
        with pytest.raises(RuntimeError, match="training workspace"):
            e.train_step(z, z, z, z, z, loss_kind="ce", ignore_index=PAD)
    finally:
        e.close()
