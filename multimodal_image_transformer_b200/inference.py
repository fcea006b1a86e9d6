"""Drop-in for the generation call and id post-processing of the reference's inference.py
(reference inference.py:17-128).  File / CLI plumbing (tokenizer files, checkpoint paths) stays the
caller's business; the two hot pieces are `generate_caption_ids` (KV-cached generation on the GPU)
and `postprocess_ids` (cut at END, drop START, strip <UNK>, squeeze spaces)."""
import re
from typing import List, Optional, Sequence

import torch

from . import config


def postprocess_ids(ids: Sequence[int], start_id: int = config.START_TOKEN_ID, end_id: int = config.END_TOKEN_ID
                    ) -> List[int]:
    """Token ids to decode: cut at the first END, drop one leading START (inference.py:98-107)."""
    ids = list(ids)
    if end_id in ids:
        ids = ids[:ids.index(end_id)]
    if ids and ids[0] == start_id:
        ids = ids[1:]
    return ids


def clean_caption(text: str, unk_token: str = config.UNK_TOKEN) -> str:
    """Strip <UNK> markers and collapse whitespace (inference.py:115-126)."""
    return re.sub(r"\s+", " ", text.replace(unk_token, "")).strip()


def generate_caption_ids(model, image, start_id: int = config.START_TOKEN_ID, end_id: int = config.END_TOKEN_ID,
                         max_len: int = config.MAX_SEQ_LEN, method: str = "greedy",
                         beam_size: int = config.BEAM_SIZE) -> List[int]:
    """model.generate(...) as inference.py:84-91 calls it."""
    return model.generate(image=image, start_token_id=start_id, end_token_id=end_id, max_len=max_len,
                          method=method, beam_size=beam_size)


def generate_caption(image_path: str, device: str, checkpoint_path: str, *, model=None, tokenizer=None,
                     method: str = "greedy") -> Optional[str]:
    """Caption for one image file (reference inference.py:17).  The caller supplies the tokenizer and
    (optionally) a constructed model; weights come from a reference-format .safetensors file."""
    from PIL import Image
    from safetensors.torch import load_file
    if tokenizer is None or model is None:
        raise ValueError("b200 generate_caption needs `model` and `tokenizer` (tokenizer training / hub download "
                         "are outside the hot path)")
    model.load_state_dict(load_file(checkpoint_path, device=str(device)))
    model.eval()
    image = Image.open(image_path).convert("RGB")
    start_id, end_id = tokenizer.token_to_id(config.START_TOKEN), tokenizer.token_to_id(config.END_TOKEN)
    ids = generate_caption_ids(model, image, start_id, end_id, config.MAX_SEQ_LEN, method)
    return clean_caption(tokenizer.decode(postprocess_ids(ids, start_id, end_id), skip_special_tokens=True))
