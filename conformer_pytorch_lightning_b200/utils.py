"""Mask helpers of the reference's src/utils.py:84-160.  Host-side integer/bool work that must be
bit exact, including the order of draws from torch's global RNG."""
import torch


def make_pad_mask(input_lengths, max_seq_len):
    """True where the frame is padding (utils.py:84-93)."""
    seq = torch.arange(0, max_seq_len, dtype=torch.int64, device=input_lengths.device)
    return seq.unsqueeze(0) >= input_lengths.unsqueeze(-1)


def subsequent_chunk_mask(size, chunk_size, num_left_chunks, device):
    """utils.py:96-111 in closed form: row i sees columns j with
    j < (i//c + 1)*c  and  (num_left_chunks < 0  or  j >= (i//c - num_left_chunks)*c)."""
    chunk_size = int(chunk_size)
    num_left_chunks = int(num_left_chunks)
    idx = torch.arange(size, device=device)
    blk = torch.div(idx, chunk_size, rounding_mode='floor')
    ret = idx.unsqueeze(0) < ((blk + 1) * chunk_size).unsqueeze(1)
    if num_left_chunks >= 0:
        ret = ret & (idx.unsqueeze(0) >= ((blk - num_left_chunks) * chunk_size).unsqueeze(1))
    return ret


def make_attn_mask(inputs, inputs_pad_mask, use_dynamic_chunk, use_dynamic_left_chunk, decoding_chunk_size,
                   static_chunk_size, num_decoding_left_chunks):
    """utils.py:115-160.  The two torch.randint draws come from the global generator in the same
    order as the reference's, in eval as well as in training."""
    max_len = inputs.size(1)
    if use_dynamic_chunk:
        if decoding_chunk_size < 0:
            chunk_size, num_left_chunks = max_len, -1
        elif decoding_chunk_size > 0:
            chunk_size, num_left_chunks = decoding_chunk_size, num_decoding_left_chunks
        else:
            chunk_size = torch.randint(1, max_len, (1,)).item()
            num_left_chunks = -1
            if chunk_size > max_len // 2:
                chunk_size = max_len
            else:
                chunk_size = chunk_size % 25 + 1
                if use_dynamic_left_chunk:
                    num_left_chunks = torch.randint(0, max_len - 1, (1,)).item()
        chunk_masks = subsequent_chunk_mask(max_len, chunk_size, num_left_chunks, inputs.device).unsqueeze(0)
        return inputs_pad_mask & chunk_masks
    if static_chunk_size > 0:
        chunk_masks = subsequent_chunk_mask(max_len, static_chunk_size, num_decoding_left_chunks,
                                            inputs.device).unsqueeze(0)
        return inputs_pad_mask & chunk_masks
    return inputs_pad_mask
