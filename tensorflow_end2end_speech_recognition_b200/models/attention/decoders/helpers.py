"""Decoding helpers -- the two ``tf.contrib.seq2seq`` helpers the reference plugs into its
attention decoder (attention_seq2seq.py:440-446 TrainingHelper, :486-494
GreedyEmbeddingHelper).  They only carry configuration; the per-step arithmetic (sample =
arg-max, embedding gather of the next input, finished flags) runs in ``b2_argmax_rows`` /
``b2_decoder_step_emit``.
"""
import torch


class GreedyEmbeddingHelper(object):
    def __init__(self, embedding, start_tokens, end_token):
        self.embedding = embedding                  # [num_classes, embedding_dim] cuda f32
        self.start_tokens = start_tokens            # [B] int32 cuda
        self.end_token = int(end_token)

    @property
    def batch_size(self):
        return int(self.start_tokens.shape[0])


class TrainingHelper(object):
    """Teacher forcing on ``labels[:, :-1]`` with ``sequence_length = labels_seq_len - 1``.

    The reference passes the already-embedded ``labels_embedded[:, :-1, :]``; this helper takes
    the label ids and the embedding table so that the gather stays inside the step kernel."""

    def __init__(self, embedding, labels, sequence_length):
        self.embedding = embedding
        self.labels = labels.to(torch.int32).contiguous()             # [B, T_out] incl. <EOS>
        self.sequence_length = sequence_length.to(torch.int32).contiguous()   # already minus one

    @property
    def batch_size(self):
        return int(self.labels.shape[0])
