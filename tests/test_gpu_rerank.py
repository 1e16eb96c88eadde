"""GPU parity of the cross-encoder engine against the float32 transformers oracle (seeded weights)."""

from __future__ import annotations

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_pairs(n, vocab, rng, lo=6, hi=180):
    ids, types = [], []
    for _ in range(n):
        L = int(rng.integers(lo, hi))
        q = int(rng.integers(2, max(3, L // 3)))
        a = rng.integers(1000, vocab, size=L).astype(np.int32)
        a[0], a[q], a[-1] = 101, 102, 102                       # [CLS] ... [SEP] ... [SEP]
        t = np.zeros(L, np.int32)
        t[q + 1:] = 1
        ids.append(a); types.append(t)
    return ids, types


def test_linear_layer_matches_torch():
    import ctypes

    import torch

    from raglite_b200 import _lib

    lib = _lib.load()
    g = torch.Generator().manual_seed(0)
    for (T, N, K, act) in [(300, 384, 384, 0), (1000, 1536, 384, 1), (777, 384, 1536, 0), (64, 1152, 384, 0)]:
        X = (torch.randn((T, K), generator=g) * 0.5).half().cuda()
        W = (torch.randn((N, K), generator=g) / K**0.5).float().cuda()
        b = torch.randn(N, generator=g).float().cuda()
        img = torch.empty(lib.rl_xenc_linear_image_bytes(N, K), dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        assert lib.rl_xenc_pack_linear(W.data_ptr(), N, K, img.data_ptr(), s) == 0
        Y = torch.empty((T, N), dtype=torch.float16, device="cuda")
        assert lib.rl_xenc_linear(X.data_ptr(), img.data_ptr(), b.data_ptr(), Y.data_ptr(), T, N, K, act, s) == 0, lib.rl_last_error()
        ref = X.float() @ W.half().float().T + b
        if act:
            ref = torch.nn.functional.gelu(ref)
        err = (Y.float() - ref).abs().max().item()
        assert err < 2e-2, (T, N, K, act, err)


@pytest.mark.parametrize("layers", [2, 12])
def test_cross_encoder_logits_match_transformers_fp32(layers):
    from scipy.stats import kendalltau

    from oracle import rerank as orr
    from raglite_b200._xenc import CrossEncoderEngine

    model = orr.seeded_model(seed=layers, num_hidden_layers=layers, vocab_size=5000)
    eng = CrossEncoderEngine.from_hf(model, max_tokens_per_call=4000)       # forces several packed calls
    rng = np.random.default_rng(1)
    ids, types = _random_pairs(48, 5000, rng)
    ids.append(np.array([101, 2000, 102, 2001, 102], np.int32)); types.append(np.array([0, 0, 0, 1, 1], np.int32))
    got_logit, got_score = eng.score_tokens(ids, types)
    want = orr.hf_logits(model, ids, types)
    assert np.abs(got_logit - want).max() < 4e-2, np.abs(got_logit - want).max()
    assert np.abs(got_score - orr.flashrank_scores(want)).max() < 1e-2
    assert kendalltau(got_logit, want)[0] > 0.97


def test_rerank_chunks_with_b200_cross_encoder(tmp_path):
    """End to end through the reference's call shape: rerank_chunks -> ranker.rank(query=, docs=)."""
    from tokenizers import Tokenizer, models, pre_tokenizers, processors

    import raglite_b200 as rl
    from oracle import rerank as orr
    from raglite_b200._rerank import ScoreFnRanker
    from raglite_b200._xenc import CrossEncoderEngine

    words = ["[PAD]", "[UNK]", "[CLS]", "[SEP]"] + [f"w{i}" for i in range(200)]
    tok = Tokenizer(models.WordPiece({w: i for i, w in enumerate(words)}, unk_token="[UNK]"))
    tok.pre_tokenizer = pre_tokenizers.Whitespace()
    tok.post_processor = processors.TemplateProcessing(single="[CLS] $A [SEP]", pair="[CLS] $A [SEP] $B:1 [SEP]:1",
                                                       special_tokens=[("[CLS]", 2), ("[SEP]", 3)])
    model = orr.seeded_model(seed=5, num_hidden_layers=3, vocab_size=len(words))
    eng = CrossEncoderEngine.from_hf(model, tok, max_length=64)
    rng = np.random.default_rng(0)
    chunks = [rl.Chunk(id=f"c{i}", body=" ".join(f"w{j}" for j in rng.integers(0, 200, size=int(rng.integers(5, 90)))))
              for i in range(20)]
    query = "w1 w2 w3 w4"
    cfg = rl.RAGLiteConfig(reranker=ScoreFnRanker(lambda q, docs: eng.score_pairs([q] * len(docs), list(docs))))
    ranked = rl.rerank_chunks(query, chunks, config=cfg)
    ids, types = eng.encode_pairs([query] * len(chunks), [str(c) for c in chunks])
    assert max(len(x) for x in ids) <= 64                                        # truncation applied
    want = orr.rank_order(orr.flashrank_scores(orr.hf_logits(model, ids, types)))
    got = [int(c.id[1:]) for c in ranked]
    assert sorted(got) == list(range(20))
    # identical order except where float32 scores are within fp16 noise of each other
    ref_scores = orr.flashrank_scores(orr.hf_logits(model, ids, types))
    for a, b in zip(got, want.tolist(), strict=True):
        assert a == b or abs(ref_scores[a] - ref_scores[b]) < 2e-2
