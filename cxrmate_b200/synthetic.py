"""Synthetic inputs for tests and bench.py: there is no
network, so no MIMIC-CXR images, no released weights and no tokenizer
vocabularies.  Everything here is an INPUT generator; no model arithmetic.

Token-id conventions pinned by the reference notebooks
(examples/cxrmate.ipynb:307-311,350-356; examples/tokenizer.ipynb:229):
[UNK]=0 [BOS]=1 [EOS]=2 [SEP]=3 [PAD]=4 [MASK]=5, [PMT]=8 [PMT-SEP]=9 [NPF]=10 [NPI]=11.
"""
from __future__ import annotations

import torch

UNK, BOS, EOS, SEP, PAD, MASK, PMT, PMT_SEP, NPF, NPI = 0, 1, 2, 3, 4, 5, 8, 9, 10, 11
N_SPECIAL = 12
DEC_VOCAB = 30000
RWD_VOCAB = 30522
RWD_PAD, RWD_UNK, RWD_CLS, RWD_SEP, RWD_MASK = 0, 100, 101, 102, 103   # BERT uncased conventions
RWD_FIRST_WORD = 1000

# special_token_ids / sections of the two SCST rollouts (scst/gen_prompt.py:209-215,282; modelling_longitudinal.py:280-282)
SPECIAL_SAMPLE, SPECIAL_GREEDY, SECTIONS = [BOS, SEP], [PMT_SEP, BOS, SEP], [0, 1, 0, 1]


def id_map() -> torch.Tensor:
    """decoder word id -> reward-model word id (1:1 word bridge between the two synthetic vocabularies)."""
    m = torch.full((DEC_VOCAB,), RWD_UNK, dtype=torch.int32)
    # injective: decoder word ids 12..29999 -> reward ids 104..30091 (above [CLS]/[SEP]/[MASK] = 101..103)
    m[N_SPECIAL:] = torch.arange(N_SPECIAL, DEC_VOCAB, dtype=torch.int32) + (RWD_MASK + 1 - N_SPECIAL)
    return m


def make_tokenizers(dec_vocab: int = DEC_VOCAB, rwd_vocab: int = RWD_VOCAB):
    """Word-level `PreTrainedTokenizerFast` stand-ins for the 30k BPE decoder tokenizer and the CXR-BERT
    WordPiece tokenizer.  Word `w<i>` is id i of the decoder vocabulary and id id_map()[i] of the reward vocabulary."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors
    from transformers import PreTrainedTokenizerFast

    specials = ["[UNK]", "[BOS]", "[EOS]", "[SEP]", "[PAD]", "[MASK]", "[X6]", "[X7]", "[PMT]", "[PMT-SEP]", "[NPF]", "[NPI]"]
    vocab = {s: i for i, s in enumerate(specials)}
    for i in range(N_SPECIAL, dec_vocab):
        vocab[f"w{i}"] = i
    tok = Tokenizer(models.WordLevel(vocab, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    dec = PreTrainedTokenizerFast(tokenizer_object=tok, unk_token="[UNK]", bos_token="[BOS]", eos_token="[EOS]",
                                  sep_token="[SEP]", pad_token="[PAD]", mask_token="[MASK]",
                                  additional_special_tokens=["[PMT]", "[PMT-SEP]", "[NPF]", "[NPI]"])

    m = id_map()
    rv = {"[PAD]": RWD_PAD, "[UNK]": RWD_UNK, "[CLS]": RWD_CLS, "[SEP]": RWD_SEP, "[MASK]": RWD_MASK}
    used = set(rv.values())
    for i in range(N_SPECIAL, dec_vocab):
        j = int(m[i])
        if j not in used:
            rv[f"w{i}"] = j
            used.add(j)
    k = 0
    for j in range(rwd_vocab):
        if j not in used:
            rv[f"[unused{k}]"] = j
            k += 1
    rt = Tokenizer(models.WordLevel(rv, unk_token="[UNK]"))
    rt.pre_tokenizer = pre_tokenizers.WhitespaceSplit()
    rt.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", special_tokens=[("[CLS]", RWD_CLS), ("[SEP]", RWD_SEP)])
    rwd = PreTrainedTokenizerFast(tokenizer_object=rt, unk_token="[UNK]", cls_token="[CLS]", sep_token="[SEP]",
                                  pad_token="[PAD]", mask_token="[MASK]")
    return dec, rwd


def make_images(B: int, N: int, size: int = 384, seed: int = 1234, n_per_study=None) -> torch.Tensor:
    """randn images; images n >= n_b of study b are exactly 0.0 (how the reference detects padding)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, N, 3, size, size, generator=g)
    if n_per_study is None:
        n_per_study = torch.randint(1, N + 1, (B,), generator=g).tolist()
    for b, n in enumerate(n_per_study):
        x[b, n:] = 0.0
    return x


def make_prompts(B: int, max_len: int = 256, seed: int = 99, vocab: int = DEC_VOCAB, no_history_every: int = 4):
    """[PMT] findings-ids [PMT-SEP] impression-ids [BOS], right padded with [PAD]; every `no_history_every`-th
    study gets the no-history prompt [PMT][NPF][PMT-SEP][NPI][BOS]."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        if no_history_every and b % no_history_every == no_history_every - 1:
            rows.append(torch.tensor([PMT, NPF, PMT_SEP, NPI, BOS]))
            continue
        L = int(torch.randint(8, max_len + 1, (1,), generator=g))
        nf = max(1, (L - 3) * 2 // 3)
        ni = max(1, L - 3 - nf)
        rows.append(torch.cat((torch.tensor([PMT]), torch.randint(N_SPECIAL, vocab, (nf,), generator=g),
                               torch.tensor([PMT_SEP]), torch.randint(N_SPECIAL, vocab, (ni,), generator=g),
                               torch.tensor([BOS]))))
    P = max(len(r) for r in rows)
    out = torch.full((B, P), PAD, dtype=torch.int64)
    for b, r in enumerate(rows):
        out[b, : len(r)] = r
    return out


def make_label_ids(B: int, min_len: int = 32, max_len: int = 256, seed: int = 7):
    """reward-model ids of the radiologist reports: [CLS] words [SEP], padded with 0; returns (ids [B,L], lens [B])."""
    g = torch.Generator().manual_seed(seed)
    lens = torch.randint(min_len, max_len + 1, (B,), generator=g)
    L = int(lens.max())
    ids = torch.zeros(B, L, dtype=torch.int64)
    for b in range(B):
        n = int(lens[b])
        ids[b, 0] = RWD_CLS
        ids[b, 1 : n - 1] = torch.randint(RWD_FIRST_WORD, RWD_VOCAB, (n - 2,), generator=g)
        ids[b, n - 1] = RWD_SEP
    return ids, lens


def ids_to_text(ids: torch.Tensor, first_word: int = N_SPECIAL) -> str:
    return " ".join(f"w{int(i)}" for i in ids if int(i) >= first_word)


# ---------------------------------------------------------------------------------------------------------------
# Throw-away SUBWORD tokenizers trained offline on a synthetic corpus: a byte-level BPE built the
# way the reference builds its decoder tokenizer (examples/tokenizer.ipynb: BPE + ByteLevel pre-tokenizer / decoder,
# specials first so that [UNK]=0 [BOS]=1 [EOS]=2 [SEP]=3 [PAD]=4 [MASK]=5 and [PMT]=8 [PMT-SEP]=9 [NPF]=10 [NPI]=11) and an
# uncased WordPiece with BERT's special ids ([PAD]=0 [UNK]=100 [CLS]=101 [SEP]=102 [MASK]=103) standing in for the
# CXR-BERT tokenizer.  With these the text round trip of the SCST step is the real one: ids -> split_and_decode_sections ->
# BPE decode -> strings -> WordPiece encode (token boundaries differ between the two vocabularies).
# ---------------------------------------------------------------------------------------------------------------
_SEED_WORDS = ("a single portable semi erect chest radiograph was obtained pulmonary aeration has decreased moderate to large "
               "layering right left pleural effusion increased loculated intra abdominal air projects over the lung base central "
               "vascular congestion is similar cardiomegaly unchanged an endotracheal tube ends cm above carina enteric passes "
               "inferiorly below film subclavian catheter terminates at cavoatrial junction no focal consolidation pneumothorax "
               "heart size normal mediastinal contours are within limits there mild bibasilar atelectasis interval improvement "
               "of edema stable appearance lines and tubes acute osseous abnormality small bilateral effusions opacity lower lobe "
               "may represent pneumonia recommend follow up comparison prior study").split()


def synthetic_corpus(n_sentences: int = 24000, seed: int = 11, n_words: int = 20000):
    """radiology-flavoured pseudo-sentences: the seed words plus n_words pronounceable pseudo-words, Zipf-like usage
    (enough distinct material for a 30k BPE / WordPiece vocabulary, so that EVERY id a random-init decoder can emit
    decodes to text)"""
    import random
    rng = random.Random(seed)
    onsets = ["b", "c", "d", "f", "g", "h", "l", "m", "n", "p", "r", "s", "t", "v", "pl", "tr", "st", "br", "ch", "th", "ph", "sc"]
    nuclei = ["a", "e", "i", "o", "u", "ae", "io", "ou", "ea"]
    codas = ["", "", "n", "r", "s", "l", "m", "x", "tic", "sis", "al", "ary", "oma", "ity", "ous", "ion"]
    words = list(dict.fromkeys(_SEED_WORDS))
    seen = set(words)
    while len(words) < n_words:
        w = "".join(rng.choice(onsets) + rng.choice(nuclei) for _ in range(rng.randint(1, 4))) + rng.choice(codas)
        if w not in seen:
            seen.add(w)
            words.append(w)
    weights = [1.0 / (1 + i) ** 0.5 for i in range(len(words))]
    out = []
    for _ in range(n_sentences):
        n = rng.randint(4, 16)
        toks = rng.choices(words, weights=weights, k=n)
        if rng.random() < 0.2:
            toks.insert(rng.randrange(n), f"{rng.randint(1, 9)}.{rng.randint(0, 9)}")
        s = " ".join(toks)
        out.append(s[0].upper() + s[1:] + ".")
    return out, words


def train_tokenizers(seed: int = 11, dec_vocab: int = DEC_VOCAB, rwd_vocab: int = RWD_VOCAB, cache_dir=None):
    """(decoder BPE tokenizer, reward-model WordPiece tokenizer) as `PreTrainedTokenizerFast`; deterministic for a seed;
    trained once per process (a few seconds) and optionally cached as tokenizer.json files under cache_dir."""
    import os

    from tokenizers import Tokenizer, decoders, models, normalizers, pre_tokenizers, processors, trainers
    from transformers import PreTrainedTokenizerFast
    key = (seed, dec_vocab, rwd_vocab)
    if key in _TOK_CACHE:
        return _TOK_CACHE[key]
    import gzip
    if cache_dir is None:      # the committed fixtures (written by this very function, ~20 s of training otherwise)
        d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "tokenizers")
        cache_dir = d if os.path.isdir(d) else None
    files = None
    if cache_dir:
        os.makedirs(cache_dir, exist_ok=True)
        files = [os.path.join(cache_dir, f"{n}_{seed}_{dec_vocab}_{rwd_vocab}.json.gz") for n in ("bpe", "wordpiece")]
    if files and all(os.path.exists(f) for f in files):
        bpe, wp = (Tokenizer.from_str(gzip.open(f, "rt", encoding="utf-8").read()) for f in files)
    else:
        corpus, _ = synthetic_corpus(seed=seed)
        bpe = Tokenizer(models.BPE(unk_token="[UNK]"))
        bpe.pre_tokenizer = pre_tokenizers.ByteLevel(add_prefix_space=False)
        bpe.decoder = decoders.ByteLevel()
        specials = ["[UNK]", "[BOS]", "[EOS]", "[SEP]", "[PAD]", "[MASK]", "[X6]", "[X7]", "[PMT]", "[PMT-SEP]", "[NPF]", "[NPI]"]
        bpe.train_from_iterator(corpus, trainers.BpeTrainer(vocab_size=dec_vocab, special_tokens=specials, show_progress=False,
                                                            initial_alphabet=pre_tokenizers.ByteLevel.alphabet()))
        wp = Tokenizer(models.WordPiece(unk_token="[UNK]"))
        wp.normalizer = normalizers.BertNormalizer(lowercase=True)
        wp.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
        wp.decoder = decoders.WordPiece()
        bert_specials = ["[PAD]"] + [f"[unused{i}]" for i in range(99)] + ["[UNK]", "[CLS]", "[SEP]", "[MASK]"]
        wp.train_from_iterator(corpus, trainers.WordPieceTrainer(vocab_size=rwd_vocab, special_tokens=bert_specials,
                                                                 show_progress=False))
        wp.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", special_tokens=[("[CLS]", RWD_CLS), ("[SEP]", RWD_SEP)])
        if files:
            for tk, f in zip((bpe, wp), files):
                with gzip.open(f, "wt", encoding="utf-8") as fh:
                    fh.write(tk.to_str())
    dec = PreTrainedTokenizerFast(tokenizer_object=bpe, unk_token="[UNK]", bos_token="[BOS]", eos_token="[EOS]",
                                  sep_token="[SEP]", pad_token="[PAD]", mask_token="[MASK]",
                                  additional_special_tokens=["[PMT]", "[PMT-SEP]", "[NPF]", "[NPI]"])
    rwd = PreTrainedTokenizerFast(tokenizer_object=wp, unk_token="[UNK]", cls_token="[CLS]", sep_token="[SEP]",
                                  pad_token="[PAD]", mask_token="[MASK]")
    assert [dec.convert_tokens_to_ids(t) for t in ("[UNK]", "[BOS]", "[EOS]", "[SEP]", "[PAD]", "[PMT]", "[PMT-SEP]", "[NPF]", "[NPI]")] == \
        [UNK, BOS, EOS, SEP, PAD, PMT, PMT_SEP, NPF, NPI]
    assert [rwd.convert_tokens_to_ids(t) for t in ("[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]")] == \
        [RWD_PAD, RWD_UNK, RWD_CLS, RWD_SEP, RWD_MASK]
    _TOK_CACHE[key] = (dec, rwd)
    return dec, rwd


_TOK_CACHE = {}
