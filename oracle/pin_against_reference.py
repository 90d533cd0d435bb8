"""Pin the oracle against the REAL reference classes and write tests/golden/*.npz.

TEST INFRASTRUCTURE.  Runs only in the authoring container, where
/root/reference is mounted (the GPU box does not have it):

    python -m oracle.pin_against_reference            # check + (re)write fixtures
    python -m oracle.pin_against_reference --check    # check only

What is compared (every check asserts; the script exits non-zero on failure):

1. the reference's `MultiCvtWithProjectionHead.forward` (imported from
   /root/reference/modules/transformers/longitudinal_model/modelling_longitudinal.py:56-90)
   vs `oracle.cvt.encode_multi`, same weights, same pixels;
   `CvtWithProjectionHead` (modelling_single.py:53-78) vs `oracle.cvt.encode_single`;
2. the reference's `LongitudinalPromptMultiCXREncoderDecoderModel.forward`
   (:173-249, with LoRA applied by the peft stand-in) vs `oracle.bert.decoder_logits`,
   teacher-forced over a padded prompt batch, with and without cache;
3. the reference's `token_ids_to_token_type_ids(_past)` (:297-364) vs the
   vectorised restatement, on random id matrices with edge cases;
4. a greedy and a sampled rollout driven through the REFERENCE forward() and
   REFERENCE token-type helpers by the Appendix-B loop vs `oracle.decode.rollout`;
6. `MultiCXREncoderDecoderModel` (modelling_multi.py:90-261) and
   `SingleCXREncoderDecoderModel` (modelling_single.py:81-249): encoder, greedy
   rollout from [BOS] through the reference forward(), teacher-forced logits
   -> tests/golden/cxrmate_ref_variants.npz;
5. structural pins from the reference notebooks: decoder parameter count
   80,769,072 + 147,456 LoRA = 80,916,528 (examples/cxrmate.ipynb:89), tied LM head,
   no-history prompt ids [8,10,9,11,1] (examples/cxrmate.ipynb:307-311).

The golden fixtures hold the REFERENCE outputs (not the oracle's).
"""
from __future__ import annotations

import argparse
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
GOLDEN = os.path.join(ROOT, "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise SystemExit("/root/reference is not mounted: pinning runs in the authoring container only")
    sys.path.insert(0, os.path.join(HERE, "_peft_standin"))
    sys.path.insert(0, REF)
    warnings.filterwarnings("ignore")
    from modules.transformers.longitudinal_model import modelling_longitudinal as ml
    from modules.transformers.single_model import modelling_single as ms
    return ml, ms


def to_reference_keys(sd):
    """oracle/weights.py names -> the names of the instantiated reference model
    (decoder wrapped by peft: SURVEY.md Appendix D last row)."""
    out = {}
    for k, v in sd.items():
        if k.startswith("decoder."):
            k2 = "decoder.base_model.model." + k[len("decoder."):]
            for nm in ("query", "key"):
                base = f"attention.self.{nm}."
                if ".attention.self." in k and f".{nm}." in k and "crossattention" not in k:
                    if "lora_A" in k or "lora_B" in k:
                        k2 = k2.replace(".weight", ".default.weight")
                    else:
                        k2 = k2.replace(base, base + "base_layer.")
            out[k2] = v
        else:
            out[k] = v
    return out


def build_reference_model(ml, sd):
    import transformers
    enc_cfg = ml.CvtWithProjectionHeadConfig(depth=[1, 4, 16], projection_size=768)
    dec_cfg = transformers.BertConfig(vocab_size=30000, num_hidden_layers=6, type_vocab_size=2, is_decoder=True,
                                      add_cross_attention=True)
    cfg = transformers.VisionEncoderDecoderConfig.from_encoder_decoder_configs(enc_cfg, dec_cfg)
    model = ml.LongitudinalPromptMultiCXREncoderDecoderModel(config=cfg)
    missing, unexpected = model.load_state_dict(to_reference_keys(sd), strict=False)
    missing = [m for m in missing if "position_ids" not in m and "token_type_ids" not in m]
    assert not missing and not unexpected, (missing[:5], unexpected[:5])
    model.eval()
    return model


def ref_rollout(model, enc, prompt_ids, special_token_ids, sections, mask_token_id, max_new, eos, pad, do_sample,
                top_k, exp_noise):
    """Appendix-B loop over the REFERENCE forward()/token-type helpers (4.41 cache semantics)."""
    ids = prompt_ids.clone()
    unfinished = torch.ones(ids.shape[0], dtype=torch.bool)
    past = None
    toks, scores = [], []
    for t in range(max_new):
        mask = (ids != mask_token_id).int()
        pos = torch.nn.functional.relu(torch.cumsum(mask, dim=1, dtype=torch.int64) - 1)
        if past is None:
            tt = model.token_ids_to_token_type_ids(ids, special_token_ids, sections)
            feed = ids
        else:
            tt = model.token_ids_to_token_type_ids_past(ids, special_token_ids, sections)
            feed, pos = ids[:, -1:], pos[:, -1:]
        out = model(encoder_outputs=enc, decoder_input_ids=feed, decoder_attention_mask=mask,
                    decoder_token_type_ids=tt, decoder_position_ids=pos, past_key_values=past, use_cache=True,
                    return_dict=True)
        past = out.past_key_values
        logits = out.logits[:, -1].float()
        if do_sample:
            from transformers.generation.logits_process import TopKLogitsWarper
            s = TopKLogitsWarper(top_k=top_k)(ids, logits)
            p = torch.softmax(s, dim=-1)
            nxt = torch.argmax(p / exp_noise[t], dim=-1)
        else:
            s = logits
            nxt = torch.argmax(s, dim=-1)
        nxt = nxt * unfinished + pad * (~unfinished)
        ids = torch.cat((ids, nxt[:, None]), dim=1)
        scores.append(s)
        unfinished = unfinished & (nxt != eos)
        if not unfinished.any():
            break
    return ids, scores


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--beam-only", action="store_true", help="write only tests/golden/cxrmate_ref_beam.npz")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    ml, ms = import_reference()
    sys.path.insert(0, ROOT)
    from oracle import bert, cvt, decode, weights

    sd = weights.make_cxrmate_weights(seed=0)
    model = build_reference_model(ml, sd)
    report = {}

    # -- 5. structural pins ------------------------------------------------
    n_dec = sum(p.numel() for p in model.decoder.parameters())
    n_lora = sum(p.numel() for n, p in model.decoder.named_parameters() if "lora_" in n)
    assert n_dec == 80916528 and n_lora == 147456, (n_dec, n_lora)
    dm = model.decoder.base_model.model
    assert dm.cls.predictions.decoder.weight.data_ptr() == dm.bert.embeddings.word_embeddings.weight.data_ptr()
    assert weights.count_params(sd) == 31532608 + 80916528, weights.count_params(sd)
    report["params"] = (n_dec, n_lora)

    # -- 1. encoder ----------------------------------------------------------
    g = torch.Generator().manual_seed(1234)
    pixels = torch.randn(2, 2, 3, 384, 384, generator=g)
    pixels[1, 1] = 0.0                                   # padded image (modelling_longitudinal.py:83)
    enc = model.encoder(pixels)
    mem_o, mask_o = cvt.encode_multi(sd, pixels)
    d_enc = (enc.last_hidden_state - mem_o).abs().max().item()
    assert torch.equal(enc.attention_mask, mask_o)
    assert d_enc < 2e-4, d_enc
    report["encoder_maxabs"] = d_enc

    single = ms.CvtWithProjectionHead(ms.CvtWithProjectionHeadConfig(depth=[1, 4, 16], projection_size=768)).eval()
    single.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")})
    s_ref = single(pixels[0]).last_hidden_state
    d_single = (s_ref - cvt.encode_single(sd, pixels[0])).abs().max().item()
    assert d_single < 2e-4, d_single
    report["encoder_single_maxabs"] = d_single

    # -- 3. token types -------------------------------------------------------
    gi = torch.Generator().manual_seed(7)
    for trial in range(50):
        L = int(torch.randint(2, 24, (1,), generator=gi))
        ids = torch.randint(0, 12, (4, L), generator=gi)
        sp = [[1, 3], [9, 1, 3], [3]][trial % 3]
        sec = [[0, 1, 0, 1][: len(sp) + 1], None][trial % 2] if len(sp) < 3 else [0, 1, 0, 1]
        a = model.token_ids_to_token_type_ids(ids, sp, sec)
        b = decode.token_type_ids_full(ids, sp, sec)
        assert torch.equal(a, b), (ids, sp, a, b)
        a = model.token_ids_to_token_type_ids_past(ids, sp, sec)
        b = decode.token_type_ids_past(ids, sp, sec)
        assert torch.equal(a, b)

    # -- 2. decoder forward, teacher forced, right-padded prompt batch --------
    PAD, BOS, EOS, SEP = 4, 1, 2, 3
    gp = torch.Generator().manual_seed(99)
    p0 = torch.cat((torch.tensor([8]), torch.randint(12, 30000, (9,), generator=gp), torch.tensor([9]),
                    torch.randint(12, 30000, (5,), generator=gp), torch.tensor([1])))
    p1 = torch.tensor([8, 10, 9, 11, 1])
    P = len(p0)
    prompt = torch.full((2, P), PAD)
    prompt[0], prompt[1, : len(p1)] = p0, p1
    tail = torch.randint(12, 30000, (2, 6), generator=gp)
    tail[0, 2] = SEP
    full = torch.cat((prompt, tail), dim=1)
    mask = (full != PAD).int()
    pos = decode.positions_from_mask(mask)
    tt = model.token_ids_to_token_type_ids(full, [9, 1, 3], [0, 1, 0, 1])
    ref_logits = model(encoder_outputs=enc, decoder_input_ids=full, decoder_attention_mask=mask,
                       decoder_token_type_ids=tt, decoder_position_ids=pos, return_dict=True).logits
    o_logits = bert.decoder_logits(sd, full, tt, pos, mask, mem_o, mask_o)
    d_dec = (ref_logits - o_logits).abs().max().item()
    assert d_dec < 5e-4, d_dec
    report["decoder_tf_maxabs"] = d_dec

    # -- 4. rollouts -----------------------------------------------------------
    T = 12
    noise = torch.empty(T, 2, 30000).exponential_(1, generator=torch.Generator().manual_seed(5))
    gold = {}
    for name, sp, smp in (("greedy", [9, 1, 3], False), ("sample", [1, 3], True)):
        r_ids, r_scores = ref_rollout(model, enc, prompt, sp, [0, 1, 0, 1], PAD, T, EOS, PAD, smp, 50, noise)
        o = decode.rollout(sd, mem_o, mask_o, prompt, special_token_ids=sp, sections=[0, 1, 0, 1], mask_token_id=PAD,
                           max_new_tokens=T, eos_token_id=EOS, pad_token_id=PAD, do_sample=smp, top_k=50,
                           exp_noise=noise)
        o_nc = decode.rollout(sd, mem_o, mask_o, prompt, special_token_ids=sp, sections=[0, 1, 0, 1],
                              mask_token_id=PAD, max_new_tokens=4, eos_token_id=EOS, pad_token_id=PAD, do_sample=smp,
                              top_k=50, exp_noise=noise, use_cache=False)
        assert torch.equal(r_ids, o.sequences), (name, r_ids[:, P:], o.sequences[:, P:])
        assert torch.equal(o_nc.sequences, o.sequences[:, : P + 4]), name
        fin = torch.isfinite(r_scores[-1])
        assert torch.equal(fin, torch.isfinite(o.scores[-1]))
        d = (r_scores[-1][fin] - o.scores[-1][fin]).abs().max().item()
        assert d < 5e-4, d
        report[f"{name}_last_score_maxabs"] = d
        report[f"{name}_min_margin"] = o.margins.min().item()
        gold[f"{name}_sequences"] = r_ids.numpy()
        gold[f"{name}_last_logits_0"] = r_scores[-1][0].numpy()
        gold[f"{name}_first_logits_1"] = r_scores[0][1].numpy()

    # -- 4b. beam search (test_step: generate(num_beams=4), gt_prompt.py:344-362): HF's _beam_search loop as restated
    #        in oracle/beam.py (pinned against transformers' generate in tests/test_beam_oracle.py) driven through the
    #        REFERENCE forward() with `past_key_values.reorder_cache(beam_idx)`, vs the same loop over the oracle decoder.
    #        EOS is biased so that hypotheses finish early and the early-stop heuristic is exercised.
    from oracle import beam as obeam
    beam_gold = {}
    dm_bias = dm.cls.predictions.bias
    for tag, eos_bias, nbeams in (("a", 8.0, 4), ("b", 10.0, 4), ("c", 9.0, 3)):
        dm_bias.data[EOS] += eos_bias
        sd_b = dict(sd)
        sd_b["decoder.cls.predictions.bias"] = sd["decoder.cls.predictions.bias"].clone()
        sd_b["decoder.cls.predictions.bias"][EOS] += eos_bias
        enc_rep = type(enc)(last_hidden_state=enc.last_hidden_state.repeat_interleave(nbeams, 0),
                            attention_mask=enc.attention_mask.repeat_interleave(nbeams, 0))
        state = {"past": None}

        def ref_step(ids, beam_idx):
            m_ = (ids != PAD).int()
            pos_ = torch.nn.functional.relu(torch.cumsum(m_, dim=1, dtype=torch.int64) - 1)
            if state["past"] is None:
                tt_, feed_ = model.token_ids_to_token_type_ids(ids, [9, 1, 3], [0, 1, 0, 1]), ids
            else:
                state["past"].reorder_cache(beam_idx)
                tt_ = model.token_ids_to_token_type_ids_past(ids, [9, 1, 3], [0, 1, 0, 1])
                feed_, pos_ = ids[:, -1:], pos_[:, -1:]
            out_ = model(encoder_outputs=enc_rep, decoder_input_ids=feed_, decoder_attention_mask=m_,
                         decoder_token_type_ids=tt_, decoder_position_ids=pos_, past_key_values=state["past"],
                         use_cache=True, return_dict=True)
            state["past"] = out_.past_key_values
            return out_.logits[:, -1].float()

        Tb = 10
        r_b = obeam.beam_search(ref_step, prompt, num_beams=nbeams, max_new_tokens=Tb, eos_token_id=EOS, pad_token_id=PAD)
        o_b = obeam.beam_rollout(sd_b, mem_o, mask_o, prompt, num_beams=nbeams, special_token_ids=[9, 1, 3],
                                 sections=[0, 1, 0, 1], mask_token_id=PAD, max_new_tokens=Tb, eos_token_id=EOS,
                                 pad_token_id=PAD)
        dm_bias.data[EOS] -= eos_bias
        assert torch.equal(r_b.sequences, o_b.sequences), (tag, r_b.sequences[:, P:], o_b.sequences[:, P:])
        assert torch.allclose(r_b.scores, o_b.scores, atol=1e-4), (tag, r_b.scores, o_b.scores)
        report[f"beam_{tag}_lengths"] = [int((row[P:] != PAD).sum()) for row in r_b.sequences]
        beam_gold[f"beam_{tag}_sequences"] = r_b.sequences.numpy()
        beam_gold[f"beam_{tag}_scores"] = r_b.scores.numpy()
        beam_gold[f"beam_{tag}_cfg"] = np.array([eos_bias, nbeams, Tb], dtype=np.float64)

    # -- 6. the prompt-free variants: MultiCXREncoderDecoderModel (modelling_multi.py:90-261) and
    #       SingleCXREncoderDecoderModel (modelling_single.py:81-249): no LoRA, prompt = [[BOS]], default sections
    #       list(range(len(special)+1)) with special_token_ids=[SEP] (multi.py:218-228, single.py:483-493), default BERT
    #       positions (arange), cross-attention mask from the encoder (multi) / none (single)
    import transformers
    from modules.transformers.multi_model import modelling_multi as mm
    sd_plain = {k: v for k, v in sd.items() if "lora_" not in k}
    var = {}
    for vname, mod, cls_name, enc_cfg_cls in (("multi", mm, "MultiCXREncoderDecoderModel", "CvtWithProjectionHeadConfig"),
                                               ("single", ms, "SingleCXREncoderDecoderModel", "CvtWithProjectionHeadConfig")):
        enc_cfg = getattr(mod, enc_cfg_cls)(depth=[1, 4, 16], projection_size=768)
        dec_cfg = transformers.BertConfig(vocab_size=30000, num_hidden_layers=6, type_vocab_size=2, is_decoder=True,
                                          add_cross_attention=True)
        cfg = transformers.VisionEncoderDecoderConfig.from_encoder_decoder_configs(enc_cfg, dec_cfg)
        vm = getattr(mod, cls_name)(config=cfg)
        missing, unexpected = vm.load_state_dict(sd_plain, strict=False)
        missing = [m for m in missing if "position_ids" not in m and "token_type_ids" not in m]
        assert not missing and not unexpected, (vname, missing[:5], unexpected[:5])
        vm.eval()
        vpx = pixels if vname == "multi" else pixels[:, 0]
        venc = vm.encoder(vpx)
        if vname == "multi":
            vmem, vmask = cvt.encode_multi(sd_plain, vpx)
            assert torch.equal(venc.attention_mask, vmask)
        else:
            vmem, vmask = cvt.encode_single(sd_plain, vpx), None
        assert (venc.last_hidden_state - vmem).abs().max().item() < 2e-4
        Tv = 8
        ids = torch.full((2, 1), BOS)
        unfinished = torch.ones(2, dtype=torch.bool)
        past, v_scores = None, []
        for t in range(Tv):                                  # Appendix-B loop over the REFERENCE forward()
            if past is None:
                tt, feed = vm.token_ids_to_token_type_ids(ids, [SEP]), ids
            else:
                tt, feed = vm.token_ids_to_token_type_ids_past(ids, [SEP]), ids[:, -1:]
            out = vm(encoder_outputs=venc, decoder_input_ids=feed, decoder_attention_mask=torch.ones_like(ids),
                     decoder_token_type_ids=tt, past_key_values=past, use_cache=True, return_dict=True)
            past = out.past_key_values
            lg = out.logits[:, -1].float()
            nxt = torch.argmax(lg, dim=-1)
            nxt = nxt * unfinished + PAD * (~unfinished)
            ids = torch.cat((ids, nxt[:, None]), dim=1)
            v_scores.append(lg)
            unfinished = unfinished & (nxt != EOS)
            if not unfinished.any():
                break
        o = decode.rollout(sd_plain, vmem, vmask, torch.full((2, 1), BOS), special_token_ids=[SEP], sections=None,
                           mask_token_id=None, max_new_tokens=Tv, eos_token_id=EOS, pad_token_id=PAD)
        assert torch.equal(ids, o.sequences), (vname, ids, o.sequences)
        d = (v_scores[-1] - o.scores[-1]).abs().max().item()
        assert d < 5e-4, (vname, d)
        report[f"{vname}_last_score_maxabs"] = d
        # teacher-forced forward of the variant on the generated ids (what model.forward returns)
        tt_full = vm.token_ids_to_token_type_ids(ids, [SEP])
        tf = vm(encoder_outputs=venc, decoder_input_ids=ids, decoder_attention_mask=torch.ones_like(ids),
                decoder_token_type_ids=tt_full, return_dict=True).logits
        var[f"{vname}_sequences"] = ids.numpy()
        var[f"{vname}_last_logits_0"] = v_scores[-1][0].numpy()
        var[f"{vname}_tf_logits_slice"] = tf[:, :, ::101].numpy()
        var[f"{vname}_tf_argmax"] = tf.argmax(-1).numpy()
        var[f"{vname}_memory_slice"] = venc.last_hidden_state[:, ::37, ::29].numpy()

    print("pin report:", report)
    if args.check:
        return
    np.savez_compressed(os.path.join(GOLDEN, "cxrmate_ref_beam.npz"), **beam_gold)
    print("wrote", os.path.join(GOLDEN, "cxrmate_ref_beam.npz"))
    if args.beam_only:
        return
    np.savez_compressed(os.path.join(GOLDEN, "cxrmate_ref_variants.npz"), T=Tv, **var)
    print("wrote", os.path.join(GOLDEN, "cxrmate_ref_variants.npz"))
    os.makedirs(GOLDEN, exist_ok=True)
    np.savez_compressed(
        os.path.join(GOLDEN, "cxrmate_ref_small.npz"),
        weights_seed=0, pixel_seed=1234, noise_seed=5, T=T, prompt=prompt.numpy(), tf_ids=full.numpy(),
        memory_slice=enc.last_hidden_state[:, ::37, ::29].numpy(), memory_mask=enc.attention_mask.numpy(),
        memory_mean=enc.last_hidden_state.mean(-1).numpy(),
        single_slice=s_ref[:, ::37, ::29].numpy(),
        tf_logits_slice=ref_logits[:, :, ::101].numpy(), tf_logits_argmax=ref_logits.argmax(-1).numpy(),
        **gold,
    )
    print("wrote", os.path.join(GOLDEN, "cxrmate_ref_small.npz"))


if __name__ == "__main__":
    main()
