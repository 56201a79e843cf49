"""Host-side helpers around the memory bank that keep the reference's function names and argument meaning:

  colab_l4_training.py:187-197  store_custom_memory          :200-206  retrieve_custom_memories
  colab_l4_training.py:209-222  one_shot_memorize_text       :225-254  one_shot_memorize_and_generate
  colab_l4_training.py:272-316  ingest_jsonl_to_memory       :319-350  ingest_csv_pairs_to_memory
  memory_augmented_layer.py:86-130  retrieve_memories (batched here: one search per query block, not per item)

The model forward and the tokenizer are the caller's (out of scope); these functions only decide what
text / features reach `create_episodic_memory` and `retrieve_similar_memories`.  They work with the reference's
`HippocampalFormation` as well as with the CUDA one; `retrieve_memories` needs the CUDA one (`retrieve_batch`).
"""
from __future__ import annotations

import csv
import json
import time
from typing import Callable, Iterable, Iterator, Optional, Tuple

import torch

# JSONL record -> text: first matching (key_a, key_b, template) wins (colab_l4_training.py:299-310)
_PAIR_FIELDS = (
    ("instruction", "output", "Instruction: {a}\nResponse: {b}"),
    ("prompt", "completion", "Prompt: {a}\nCompletion: {b}"),
    ("input", "output", "Input: {a}\nOutput: {b}"),
)


def store_custom_memory(hippocampus, features: torch.Tensor, memory_id: Optional[str] = None):
    """Write an external feature vector (2-D inputs are mean-pooled over dim 0) into the bank; returns the id."""
    if hippocampus is None:
        return None
    vec = features.detach()
    if vec.dim() == 2:
        vec = vec.mean(dim=0)
    mid = memory_id if memory_id is not None else f"external-{int(time.time())}"
    hippocampus.create_episodic_memory(memory_id=mid, event_id=mid, features=vec)
    return mid


def retrieve_custom_memories(hippocampus, query_features: torch.Tensor, location: Optional[torch.Tensor] = None,
                             k: int = 5):
    """Top-k (memory_id, score) for a query vector; multi-row queries are mean-pooled first."""
    if hippocampus is None:
        return []
    q = query_features.mean(dim=0) if query_features.dim() > 1 else query_features
    return hippocampus.retrieve_similar_memories(q, location=location, k=k)


def one_shot_memorize_text(text: str, tokenizer, model, hippocampus, device, memory_id: Optional[str] = None):
    """Encode `text`, run the model once with store_memory=True so that its forward writes the pooled hidden
    state into the bank under `memory_id` (hippocampal_transformer.py:125-138)."""
    if hippocampus is None or model is None or tokenizer is None:
        return None
    limit = getattr(getattr(model, "config", None), "max_seq_len", 256)
    ids = tokenizer.encode(text, return_tensors='pt', truncation=True, max_length=limit).to(device)
    mid = memory_id or f"oneshot-{int(time.time())}"
    model.eval()
    with torch.no_grad():
        model(ids, prosody=None, use_memory=False, store_memory=True, memory_ids=[mid])
    return mid


def one_shot_memorize_and_generate(support_text: str, prompt: str, tokenizer, model, hippocampus, device,
                                   max_new_tokens: int = 40, temperature: float = 0.7) -> str:
    """Store `support_text`, then sample a continuation of `prompt` with memory retrieval switched on."""
    one_shot_memorize_text(support_text, tokenizer, model, hippocampus, device)
    window = getattr(getattr(model, "config", None), "max_seq_len", 256)
    model.eval()
    tokens = tokenizer.encode(prompt, return_tensors='pt').to(device)
    eos = getattr(tokenizer, "eos_token_id", None)
    with torch.no_grad():
        for _ in range(max_new_tokens):
            logits, _ = model(tokens[:, -window:], use_memory=True, store_memory=False)
            probs = torch.softmax(logits[:, -1, :] / temperature, dim=-1)
            nxt = torch.multinomial(probs, num_samples=1)
            tokens = torch.cat([tokens, nxt], dim=1)
            if eos is not None and bool((nxt == eos).all()):
                break
    return tokenizer.decode(tokens[0].tolist(), skip_special_tokens=True)


def _jsonl_texts(path: str) -> Iterator[str]:
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for raw in fh:
            raw = raw.strip()
            if not raw:
                continue
            try:
                rec = json.loads(raw)
            except Exception:
                continue
            text = None
            if isinstance(rec, str):
                text = rec
            elif isinstance(rec, dict):
                if "text" in rec:
                    text = rec["text"]
                else:
                    for ka, kb, tmpl in _PAIR_FIELDS:
                        if ka in rec and kb in rec:
                            text = tmpl.format(a=rec[ka], b=rec.get(kb, ''))
                            break
            if text:
                yield text


def _csv_texts(path: str, delimiter: str) -> Iterator[str]:
    with open(path, "r", encoding="utf-8", errors="ignore") as fh:
        for row in csv.reader(fh, delimiter=delimiter):
            if len(row) < 2:
                continue
            q, a = row[0].strip(), row[1].strip()
            if q or a:
                yield f"Question: {q}\nAnswer: {a}"


def _ingest(texts: Iterable[str], prefix: str, tokenizer, model, hippocampus, device, max_items: int,
            memorize: Optional[Callable]) -> int:
    if hippocampus is None or model is None or tokenizer is None:
        return 0
    memorize = memorize or one_shot_memorize_text
    stored = 0
    for text in texts:
        if stored >= max_items:
            break
        memorize(text, tokenizer, model, hippocampus, device, memory_id=f"{prefix}-{stored}")
        stored += 1
    return stored


def ingest_jsonl_to_memory(path: str, tokenizer, model, hippocampus, device, max_items: int = 1000,
                           memorize: Optional[Callable] = None) -> int:
    """Stream a JSONL file into episodic memory (ids `jsonl-<n>`); returns the number stored.  Lines may be a JSON
    string, {"text": ...} or an (instruction/output | prompt/completion | input/output) pair."""
    return _ingest(_jsonl_texts(path), "jsonl", tokenizer, model, hippocampus, device, max_items, memorize)


def ingest_csv_pairs_to_memory(path: str, tokenizer, model, hippocampus, device, max_items: int = 1000,
                               delimiter: str = ",", memorize: Optional[Callable] = None) -> int:
    """Stream a two-column CSV (question, answer) into episodic memory (ids `csv-<n>`)."""
    return _ingest(_csv_texts(path, delimiter), "csv", tokenizer, model, hippocampus, device, max_items, memorize)


def ingest_features_to_memory(hippocampus, features: torch.Tensor, prefix: str = "bulk") -> int:
    """Bulk form for pre-computed embeddings [N,d]: one `create_episodic_memories` call."""
    n = features.shape[0]
    first = hippocampus.memory_count
    hippocampus.create_episodic_memories(features, [f"{prefix}-{first + i}" for i in range(n)])
    return n


def retrieve_memories(hippocampus, query: torch.Tensor, k: int = 5) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched body of MemoryAugmentedLayer.retrieve_memories (memory_augmented_layer.py:108-130): query [B,D]
    (the layer's `query_proj(hidden.mean(1))`) -> (memory_features [B,k,D], memory_scores [B,k]), zero-padded
    where fewer than k memories exist, in the query's dtype.  One search per block instead of one per item."""
    b, d = query.shape
    feats = torch.zeros(b, k, d, device=query.device, dtype=query.dtype)
    scores = torch.zeros(b, k, device=query.device, dtype=query.dtype)
    if hippocampus is None or hippocampus.memory_count == 0:
        return feats, scores
    idx, sc, rows = hippocampus.retrieve_batch(query.detach().float(), k=k, gather=True)
    kk = idx.shape[1]
    ok = (idx >= 0).unsqueeze(-1)
    feats[:, :kk] = torch.where(ok, rows, torch.zeros_like(rows)).to(query.dtype)
    scores[:, :kk] = torch.where(ok.squeeze(-1), sc, torch.zeros_like(sc)).to(query.dtype)
    return feats, scores
