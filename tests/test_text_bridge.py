"""The text bridge with SUBWORD tokenizers (byte-level BPE decoder vocabulary, uncased WordPiece reward vocabulary; trained
offline by cxrmate_b200.synthetic.train_tokenizers and committed under tests/golden/tokenizers): the product's
split / decode / encode path and prompt tokenisation against the oracle's restatement of the reference methods
(modelling_longitudinal.py:413-457,459-513; tools/rewards/cxrbert.py:33-40)."""
import pytest
import torch

PAD, BOS, EOS, SEP, PMT, PMT_SEP = 4, 1, 2, 3, 8, 9


@pytest.fixture(scope="module")
def toks():
    from cxrmate_b200 import synthetic as S
    return S.train_tokenizers()


def _rows(dec, reports, L):
    rows = []
    for prompt, f, i, tail in reports:
        ids = list(prompt)
        ids += dec(f, add_special_tokens=False)["input_ids"] if f is not None else []
        if i is not None:
            ids += [SEP] + dec(i, add_special_tokens=False)["input_ids"]
        ids += tail
        assert len(ids) <= L
        rows.append(ids + [PAD] * (L - len(ids)))
    return torch.tensor(rows)


def test_bridge_round_trip_and_oracle_agreement(toks):
    from cxrmate_b200.modelling import CXRMateEngineModel
    from cxrmate_b200.text_bridge import TextBridge
    from oracle import text
    dec, rwd = toks
    prompt = [PMT, 10, PMT_SEP, 11, BOS]
    f1, i1 = "Moderate right pleural effusion has increased.", "No pneumothorax."
    f2 = "Heart size is normal. Lines and tubes unchanged 2.5 cm above carina."
    seq = _rows(dec, [(prompt, f1, i1, [EOS]), (prompt, f2, None, []), (prompt, None, "", [EOS]), (prompt, f1, i1, [])], 48)
    br = TextBridge(dec, rwd, BOS, SEP, EOS)
    texts, ids, lens = br(seq)
    assert texts[0] == f"{f1} {i1}" and texts[1] == f"{f2} " and texts[2] == " "      # byte-level BPE decodes losslessly
    # the reference's per-row path, restated by the oracle, and the model class's method give the same strings
    _, f, i = text.split_and_decode_sections(seq, [BOS, SEP, EOS], dec)
    assert texts == [f"{a} {b}" for a, b in zip(f, i)]
    m = CXRMateEngineModel.__new__(CXRMateEngineModel)
    assert m.split_and_decode_sections(seq, [BOS, SEP, EOS], dec) == text.split_and_decode_sections(seq, [BOS, SEP, EOS], dec)
    # reward ids: exactly CXRBERTReward's tokenisation of those strings
    want = rwd(texts, add_special_tokens=True, padding="longest", return_tensors="pt", truncation=True, max_length=512)
    assert torch.equal(ids.long(), want["input_ids"]) and torch.equal(lens.long(), want["attention_mask"].sum(1))
    assert int(ids[0, 0]) == 101 and int(ids[0, int(lens[0]) - 1]) == 102
    # the two vocabularies split the same text differently: this is a real re-tokenisation, not an id map
    n_bpe = len(dec(f1, add_special_tokens=False)["input_ids"])
    n_wp = len(rwd(f1, add_special_tokens=False)["input_ids"])
    assert n_bpe != n_wp or dec.tokenize(f1) != rwd.tokenize(f1)


def test_bridge_random_ids_truncate_at_512(toks):
    """ids of a random-init decoder: every id decodes, the reward input is truncated at max_position_embeddings"""
    from cxrmate_b200.text_bridge import TextBridge
    dec, rwd = toks
    g = torch.Generator().manual_seed(3)
    seq = torch.cat((torch.tensor([[PMT, 10, PMT_SEP, 11, BOS]] * 4), torch.randint(12, 30000, (4, 255), generator=g)), 1)
    texts, ids, lens = TextBridge(dec, rwd, BOS, SEP, EOS)(seq)
    assert all(len(t) > 100 for t in texts)
    assert ids.shape[1] <= 512 and int(lens.max()) <= 512 and int(ids.max()) < 30522 and int(ids.min()) >= 0
    # the bridge calls the Rust tokenizers' batch routines directly: same strings and ids as the transformers wrappers the
    # reference goes through (tokenizer.decode per section, batch_encode_plus with padding='longest' and truncation)
    br = TextBridge(dec, rwd, BOS, SEP, EOS)
    f, i = br.split_ids(seq)
    assert texts == [f"{a} {b}" for a, b in zip(dec.batch_decode(f, skip_special_tokens=True), dec.batch_decode(i, skip_special_tokens=True))]
    want = rwd(texts, add_special_tokens=True, padding="longest", return_tensors="pt", truncation=True, max_length=512)
    assert torch.equal(ids.long(), want["input_ids"]) and torch.equal(lens.long(), want["attention_mask"].sum(1))
    short = TextBridge(dec, rwd, BOS, SEP, EOS, max_reward_len=64)(seq)      # truncation inside the batch call
    want64 = rwd(texts, add_special_tokens=True, padding="longest", return_tensors="pt", truncation=True, max_length=64)
    assert torch.equal(short[1].long(), want64["input_ids"]) and int(short[2].max()) == 64


def test_tokenize_prompt_and_report_with_bpe(toks):
    from cxrmate_b200.modelling import CXRMateEngineModel
    from oracle import text
    dec, _ = toks
    m = CXRMateEngineModel.__new__(CXRMateEngineModel)
    m.device = torch.device("cpu")
    pf = ["Moderate right pleural effusion.", None, "Heart size is normal. " * 40]
    pi = ["No pneumothorax.", None, "Stable."]
    for max_len in (256, 24):
        a = m.tokenize_prompt(pf, pi, dec, max_len, add_bos_token_id=True)
        b = text.tokenize_prompt(pf, pi, dec, max_len, add_bos_token_id=True)
        assert torch.equal(a["input_ids"], b["input_ids"]) and torch.equal(a["attention_mask"], b["attention_mask"])
        assert a["input_ids"].shape[1] <= max_len
        assert a["input_ids"][1, :5].tolist() == [8, 10, 9, 11, 1]                      # no-history prompt
        assert int(a["input_ids"][2, -1]) == BOS                                        # truncated row: last column forced to BOS
    tf = m.tokenize_report_teacher_forcing(["Moderate effusion."], ["No change."], dec, 32)
    full = dec("[BOS]Moderate effusion.[SEP]No change.[EOS]", add_special_tokens=False)["input_ids"]
    assert tf["decoder_input_ids"][0].tolist() == full[:-1] and tf["label_ids"][0].tolist() == full[1:]
