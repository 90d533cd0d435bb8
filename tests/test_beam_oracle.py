"""oracle/beam.py (the restated beam search) pinned against `transformers` `generate(num_beams=...)` itself, on a small
HF decoder where the installed generate() is sound (a plain GPT-2 without the reference's prepare_inputs override), and
self-consistency of the cxrmate step function (cached + reordered == uncached)."""
import pytest
import torch


def _tiny_gpt2(seed, vocab=40):
    from transformers import GPT2Config, GPT2LMHeadModel
    torch.manual_seed(seed)
    cfg = GPT2Config(vocab_size=vocab, n_positions=64, n_embd=32, n_layer=2, n_head=2, bos_token_id=1, eos_token_id=2,
                     pad_token_id=0)
    m = GPT2LMHeadModel(cfg).eval()
    with torch.no_grad():
        for p in m.parameters():          # larger weights: peaked distributions, early EOS, real beam competition
            p.mul_(4.0)
    return m


@pytest.mark.parametrize("seed,nb,T,lp", [(0, 4, 12, 1.0), (1, 4, 20, 1.0), (2, 3, 16, 1.0), (3, 4, 16, 2.0), (4, 2, 10, 0.5),
                                          (5, 4, 24, 1.0), (6, 5, 18, 1.0)])
def test_beam_loop_matches_hf_generate(seed, nb, T, lp):
    from oracle.beam import beam_search
    m = _tiny_gpt2(seed)
    g = torch.Generator().manual_seed(100 + seed)
    B, P = 5, 4
    prompt = torch.randint(3, 40, (B, P), generator=g)
    with torch.no_grad():
        ref = m.generate(input_ids=prompt, attention_mask=torch.ones_like(prompt), num_beams=nb, max_new_tokens=T,
                         do_sample=False, length_penalty=lp, early_stopping=False, eos_token_id=2, pad_token_id=0,
                         return_dict_in_generate=True, output_scores=True, use_cache=False)

        def step(ids, beam_idx):
            return m(input_ids=ids).logits[:, -1]

        out = beam_search(step, prompt, num_beams=nb, max_new_tokens=T, eos_token_id=2, pad_token_id=0, length_penalty=lp)
    assert out.sequences.shape == ref.sequences.shape, (out.sequences.shape, ref.sequences.shape)
    assert torch.equal(out.sequences, ref.sequences)
    assert torch.allclose(out.scores, ref.sequences_scores, atol=1e-5)


def test_cxrmate_beam_cached_equals_uncached():
    """the oracle's cached step function (self K/V reordered by beam_idx) == recomputing every running beam from scratch"""
    from cxrmate_b200 import synthetic as S
    from oracle import bert, weights
    from oracle.beam import beam_rollout, beam_search
    from oracle.decode import positions_from_mask, token_type_ids_full
    sd = weights.make_cxrmate_weights(seed=0)
    g = torch.Generator().manual_seed(5)
    B, nb, T = 2, 3, 6
    mem = torch.randn(B, 7, 768, generator=g)
    mmask = torch.ones(B, 7, dtype=torch.bool)
    mmask[1, 4:] = False
    prompt = S.make_prompts(B, 8, seed=3)
    kw = dict(num_beams=nb, special_token_ids=S.SPECIAL_GREEDY, sections=S.SECTIONS, mask_token_id=S.PAD, layers=2)
    a = beam_rollout(sd, mem, mmask, prompt, max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD, **kw)

    def step(ids, beam_idx):
        mask = (ids != S.PAD).long()
        tt = token_type_ids_full(ids, S.SPECIAL_GREEDY, S.SECTIONS)
        # the `_past` rule types the last token by what precedes it; the full rule agrees except when the last token
        # is itself the first special token - recompute that column the cached way
        from oracle.decode import token_type_ids_past
        tt[:, -1:] = token_type_ids_past(ids, S.SPECIAL_GREEDY, S.SECTIONS) if ids.shape[1] > prompt.shape[1] else tt[:, -1:]
        return bert.decoder_logits(sd, ids, tt, positions_from_mask(mask), mask, mem.repeat_interleave(nb, 0),
                                   mmask.repeat_interleave(nb, 0), None, 2, last_only=True)[:, -1]

    with torch.no_grad():
        b = beam_search(step, prompt, num_beams=nb, max_new_tokens=T, eos_token_id=S.EOS, pad_token_id=S.PAD)
    assert torch.equal(a.sequences, b.sequences)
    assert torch.allclose(a.scores, b.scores, atol=1e-4)
