"""CPU suite: the oracle against the golden fixtures written from the REAL reference classes
(oracle/pin_against_reference.py), the structural pins of the reference notebooks, and the
properties SURVEY.md section 4 lists (cached == uncached, `_past` types == full types,
multinomial == exponential-race argmax)."""
import os

import numpy as np
import pytest
import torch

from oracle import bert, cvt, decode, scst, text, weights

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cxrmate_ref_small.npz")
PAD, BOS, EOS, SEP, PMT_SEP = 4, 1, 2, 3, 9


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


@pytest.fixture(scope="module")
def sd():
    return weights.make_cxrmate_weights(seed=0)


@pytest.fixture(scope="module")
def enc(sd):
    g = torch.Generator().manual_seed(1234)
    px = torch.randn(2, 2, 3, 384, 384, generator=g)
    px[1, 1] = 0.0
    with torch.no_grad():
        return cvt.encode_multi(sd, px) + (px,)


def test_structural_pins(sd):
    # examples/cxrmate.ipynb:89 "trainable params: 147456 || all params: 80916528"; encoder 31,532,608 (SURVEY finding 3)
    dec = {k: v for k, v in sd.items() if k.startswith("decoder.")}
    enc_ = {k: v for k, v in sd.items() if k.startswith("encoder.")}
    assert weights.count_params(dec) == 80916528
    assert sum(v.numel() for k, v in dec.items() if "lora_" in k) == 147456
    assert weights.count_params(enc_) == 31532608
    assert sd["decoder.cls.predictions.decoder.weight"].data_ptr() == sd["decoder.bert.embeddings.word_embeddings.weight"].data_ptr()


def test_encoder_matches_reference_golden(enc, gold):
    mem, mask, _ = enc
    assert np.array_equal(mask.numpy(), gold["memory_mask"])
    assert np.abs(mem[:, ::37, ::29].numpy() - gold["memory_slice"]).max() < 1e-5
    assert np.abs(mem.mean(-1).numpy() - gold["memory_mean"]).max() < 1e-5


def test_single_encoder_matches_reference_golden(sd, enc, gold):
    with torch.no_grad():
        s = cvt.encode_single(sd, enc[2][0])
    assert np.abs(s[:, ::37, ::29].numpy() - gold["single_slice"]).max() < 1e-5


def test_teacher_forced_logits_match_reference_golden(sd, enc, gold):
    ids = torch.from_numpy(gold["tf_ids"])
    mask = (ids != PAD).int()
    with torch.no_grad():
        logits = bert.decoder_logits(sd, ids, decode.token_type_ids_full(ids, [PMT_SEP, BOS, SEP], [0, 1, 0, 1]),
                                     decode.positions_from_mask(mask), mask, enc[0], enc[1])
    assert np.abs(logits[:, :, ::101].numpy() - gold["tf_logits_slice"]).max() < 1e-4
    assert np.array_equal(logits.argmax(-1).numpy(), gold["tf_logits_argmax"])


@pytest.mark.parametrize("mode", ["greedy", "sample"])
def test_rollouts_match_reference_golden(sd, enc, gold, mode):
    T = int(gold["T"])
    noise = torch.empty(T, 2, 30000).exponential_(1, generator=torch.Generator().manual_seed(int(gold["noise_seed"])))
    sp = [PMT_SEP, BOS, SEP] if mode == "greedy" else [BOS, SEP]
    with torch.no_grad():
        o = decode.rollout(sd, enc[0], enc[1], torch.from_numpy(gold["prompt"]), special_token_ids=sp,
                           sections=[0, 1, 0, 1], mask_token_id=PAD, max_new_tokens=T, eos_token_id=EOS,
                           pad_token_id=PAD, do_sample=mode == "sample", top_k=50, exp_noise=noise)
    assert np.array_equal(o.sequences.numpy(), gold[f"{mode}_sequences"])
    ref = torch.from_numpy(gold[f"{mode}_last_logits_0"])
    fin = torch.isfinite(ref)
    assert torch.equal(torch.isfinite(o.scores[-1][0]), fin)
    assert (o.scores[-1][0][fin] - ref[fin]).abs().max() < 1e-4


def test_cached_equals_uncached(sd, enc, gold):
    kw = dict(special_token_ids=[PMT_SEP, BOS, SEP], sections=[0, 1, 0, 1], mask_token_id=PAD, max_new_tokens=3,
              eos_token_id=EOS, pad_token_id=PAD)
    with torch.no_grad():
        a = decode.rollout(sd, enc[0], enc[1], torch.from_numpy(gold["prompt"]), **kw)
        b = decode.rollout(sd, enc[0], enc[1], torch.from_numpy(gold["prompt"]), use_cache=False, **kw)
    assert torch.equal(a.sequences, b.sequences)
    assert (a.scores[-1] - b.scores[-1]).abs().max() < 1e-4


def test_past_types_equal_last_column_of_full_types():
    g = torch.Generator().manual_seed(3)
    for _ in range(200):
        L = int(torch.randint(3, 30, (1,), generator=g))
        ids = torch.randint(0, 14, (5, L), generator=g)
        ids[:, 0] = 8                                   # column 0 is [PMT] on the real path
        for sp in ([BOS, SEP], [PMT_SEP, BOS, SEP]):
            full = decode.token_type_ids_full(ids, sp, [0, 1, 0, 1])
            past = decode.token_type_ids_past(ids, sp, [0, 1, 0, 1])
            # equal except when two listed specials appear in reversed order, where the reference's two rules differ
            for r in range(5):
                firsts = [int((ids[r, :-1] == t).int().argmax()) if (ids[r, :-1] == t).any() else -1 for t in sp]
                present = [f for f in firsts if f >= 0]
                if present == sorted(present):
                    assert full[r, -1] == past[r, 0]


def test_token_types_vs_reference_double_loop():
    """literal transcription of the loop semantics (modelling_longitudinal.py:315-336) on random ids"""
    g = torch.Generator().manual_seed(5)
    for _ in range(100):
        L = int(torch.randint(2, 20, (1,), generator=g))
        ids = torch.randint(0, 12, (3, L), generator=g)
        sp, sec = [9, 1, 3], [0, 1, 0, 1]
        want = torch.full_like(ids, sec[0])
        for i, tok in enumerate(sp):
            for r in range(ids.shape[0]):
                col = int((ids[r] == tok).int().argmax()) + 1
                if col != 1 and col < L:
                    want[r, col:] = sec[i + 1]
        assert torch.equal(decode.token_type_ids_full(ids, sp, sec), want)


def test_product_helpers_equal_oracle_helpers():
    from cxrmate_b200 import modelling as M
    g = torch.Generator().manual_seed(9)
    for _ in range(50):
        ids = torch.randint(0, 14, (4, int(torch.randint(2, 25, (1,), generator=g))), generator=g)
        for sp, sec in (([1, 3], [0, 1, 0, 1]), ([9, 1, 3], [0, 1, 0, 1]), ([3], None)):
            assert torch.equal(M.token_ids_to_token_type_ids(ids, sp, sec), decode.token_type_ids_full(ids, sp, sec))
            assert torch.equal(M.token_ids_to_token_type_ids_past(ids, sp, sec), decode.token_type_ids_past(ids, sp, sec))
        m = (ids != 4).int()
        assert torch.equal(M.position_ids_from_mask(m), decode.positions_from_mask(m))


def test_multinomial_is_exponential_race_argmax():
    """SURVEY finding 6: torch.multinomial(p, 1) == argmax(p / q), q = empty_like(p).exponential_(1), same generator"""
    p = torch.softmax(torch.randn(6, 30000, generator=torch.Generator().manual_seed(1)) * 3, -1)
    a = torch.multinomial(p, 1, generator=torch.Generator().manual_seed(77))[:, 0]
    q = torch.empty_like(p).exponential_(1, generator=torch.Generator().manual_seed(77))
    assert torch.equal(a, torch.argmax(p / q, -1))


def test_top_k_mask_keeps_ties():
    s = torch.tensor([[5.0, 1.0, 3.0, 3.0, 0.0]])
    m = decode.top_k_mask(s, 2)
    assert torch.isfinite(m).tolist() == [[True, False, True, True, False]]


def test_reinforce_loss_formula():
    g = torch.Generator().manual_seed(2)
    B, V, T = 3, 50, 5
    logits = decode.top_k_mask(torch.randn(B, T, V, generator=g), 10).transpose(1, 2)      # [B,V,T] like torch.stack(scores,-1)
    ids = torch.stack([torch.isfinite(logits[b, :, t]).nonzero()[0, 0] for b in range(B) for t in range(T)]).view(B, T)
    ids[1, 3:] = PAD
    adv = torch.tensor([0.3, -0.2, 0.1])
    want = 0.0
    for b in range(B):
        lp = 0.0
        for t in range(T):
            if ids[b, t] != PAD:
                lp += logits[b, ids[b, t], t] - torch.logsumexp(logits[b, :, t], 0)
        want += -lp * adv[b]
    assert torch.allclose(scst.reinforce_loss(logits, ids, adv), want / B, atol=1e-5)


def test_text_bridge_and_prompt_golden():
    from cxrmate_b200 import synthetic as S
    from cxrmate_b200.modelling import CXRMateEngineModel
    dec, rwd = S.make_tokenizers()
    # examples/cxrmate.ipynb:307-311: tokenize_prompt(None, None) -> [[8, 10, 9, 11, 1]], mask all ones
    out = text.tokenize_prompt([None], [None], dec, 256, add_bos_token_id=True)
    assert out["input_ids"].tolist() == [[8, 10, 9, 11, 1]] and out["attention_mask"].tolist() == [[1] * 5]
    out = text.tokenize_prompt(["w20 w21", None], ["w30", None], dec, 6, add_bos_token_id=True)
    assert out["input_ids"].tolist() == [[8, 20, 21, 9, 30, 1], [8, 10, 9, 11, 1, 4]]
    out = text.tokenize_prompt(["w20 w21 w22 w23"], ["w30"], dec, 6, add_bos_token_id=True)     # truncated: BOS forced
    assert out["input_ids"].tolist() == [[8, 20, 21, 22, 23, 1]]
    ids = torch.tensor([[8, 10, 9, 11, 1, 50, 51, 3, 60, 2, 4, 4],
                        [8, 40, 9, 41, 1, 70, 71, 72, 73, 74, 75, 76],       # no SEP / EOS: findings run to the end
                        [8, 40, 9, 41, 1, 3, 2, 4, 4, 4, 4, 4]])            # empty sections
    a = text.split_and_decode_sections(ids, [BOS, SEP, EOS], dec)
    assert a[1] == ["w50 w51", "w70 w71 w72 w73 w74 w75 w76", ""] and a[2] == ["w60", "", ""]
    m = CXRMateEngineModel.__new__(CXRMateEngineModel)
    assert m.split_and_decode_sections(ids, [BOS, SEP, EOS], dec) == a
    # word bridge: decoder word -> reward id is what id_map says
    enc_ids = rwd(["w50 w51 w60"], return_tensors="pt")["input_ids"][0].tolist()
    assert enc_ids == [S.RWD_CLS] + [int(S.id_map()[i]) for i in (50, 51, 60)] + [S.RWD_SEP]


def test_reward_is_cosine_of_cls_projections():
    rsd = weights.make_cxrbert_weights(seed=1, layers=2)
    from oracle import reward
    g = torch.Generator().manual_seed(4)
    ids = torch.randint(1000, 30522, (3, 12), generator=g)
    ids[:, 0] = 101
    m = torch.ones_like(ids)
    with torch.no_grad():
        r = reward.reward_from_ids(rsd, ids, m, ids, m, layers=2)
        assert torch.allclose(r, torch.ones(3), atol=1e-6)
        # padding does not change the embedding
        e1 = bert.cxrbert_cls_projection(rsd, ids[:, :8], m[:, :8], 2)
        ids2 = torch.cat((ids[:, :8], torch.zeros(3, 4, dtype=torch.long)), 1)
        m2 = torch.cat((m[:, :8], torch.zeros(3, 4, dtype=torch.long)), 1)
        e2 = bert.cxrbert_cls_projection(rsd, ids2, m2, 2)
        assert (e1 - e2).abs().max() < 1e-5
