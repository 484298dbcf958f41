"""Data2Seq dispatcher drop-in (Data2Seq/Data2Seq.py:19-55): `Data2Seq(modality, dim)(data) -> (B, N, dim)`.

Deviations from the file as shipped (it cannot be imported without `clip`, and several branches crash —
SURVEY.md §8a row T0); the constructor / forward signature and attribute names (`modality`, `embed_dim`, `embed`)
are kept and the *intended* semantics implemented:
  image / infrared / x-ray : Image.PatchEmbed(embed_dim=dim)                              (Data2Seq.py:25-26)
  video                    : Video.PatchEmbed(embed_dim=dim), VideoMAE tubelets           (Data2Seq.py:29-30)
  audio                    : the reference has no `embed` for audio and returns a raw fbank (Data2Seq.py:38-45,53-54);
                             here a (B, T, F) fbank is tokenised like AST (ast_models.py:86,152-155) with
                             Acoustic.PatchEmbed(embed_dim=dim, fstride=10, tstride=10)
  time-series / imu        : Time_Series.DataEmbedding(c_in=1, d_model=dim)                  (Data2Seq.py:35-36, with the
                             `cin=` typo fixed)
  hyper                    : Hyper_Spectrum.PatchEmbed(embed_dim=dim): Linear + cls + pos (Data2Seq.py:33-34; working
                             form Hyper-spectrum/metatransformer.py:152-159)
  text                     : the reference constructs a CLIPTokenizer that returns ids, not embeddings, and its padding
                             branch is dead code (Data2Seq.py:27-28,50-52); here pre-computed (B, c) sentence features
                             are zero-padded to `dim` as Text.zero_padding does (Text.py:15-19) -> (B, 1, dim)
  graph                    : Graph.GraphFeatureTokenizer with hidden_dim = dim (the shipped call leaves hidden_dim = 1,
                             Data2Seq.py:31-32): forward(batched_data) returns the (B, 2+T, dim) tokens; the padding mask /
                             padded index the reference also returns are available from `self.embed(batched_data)`.  The
                             reference's own encoder loop feeds the tokens to the Blocks WITHOUT a key-padding mask
                             (Graph/metatransformer/modules/tokengt_graph_encoder.py:330-334), so neither does this path
"""
import torch.nn as nn

from . import Acoustic, Graph, Hyper_Spectrum, Image, Text, Time_Series, Video

_IMAGE_LIKE = ('image', 'infrared', 'x-ray')


class Data2Seq(nn.Module):

    def __init__(self, modality, dim, token_dtype=None):
        super().__init__()
        self.modality = modality
        self.embed_dim = dim
        if self.modality in _IMAGE_LIKE:
            self.embed = Image.PatchEmbed(embed_dim=self.embed_dim, token_dtype=token_dtype)
        elif self.modality == 'video':
            self.embed = Video.PatchEmbed(embed_dim=self.embed_dim, token_dtype=token_dtype)
        elif self.modality == 'audio':
            self.embed = Acoustic.PatchEmbed(embed_dim=self.embed_dim, fstride=10, tstride=10, token_dtype=token_dtype)
        elif self.modality == 'hyper':
            self.embed = Hyper_Spectrum.PatchEmbed(embed_dim=self.embed_dim, token_dtype=token_dtype)
        elif self.modality in ('time-series', 'imu'):
            self.embed = Time_Series.DataEmbedding(c_in=1, d_model=self.embed_dim, token_dtype=token_dtype)
        elif self.modality == 'text':
            self.embed = Text.TextFeaturePad(embed_dim=self.embed_dim)
        elif self.modality == 'graph':
            # TokenGT's PCQM4Mv2 vocabulary (Graph/metatransformer/models/tokengt.py defaults) with hidden_dim = dim
            self.embed = Graph.GraphFeatureTokenizer(num_atoms=512 * 9, num_edges=512 * 3, rand_node_id=False,
                                                     rand_node_id_dim=64, orf_node_id=True, orf_node_id_dim=64,
                                                     lap_node_id=False, lap_node_id_k=16, lap_node_id_sign_flip=False,
                                                     lap_node_id_eig_dropout=0.0, type_id=True, hidden_dim=self.embed_dim,
                                                     n_layers=12, token_dtype=token_dtype)
        else:
            raise ValueError(f"unknown modality '{modality}'")

    def forward(self, data):
        if self.modality == 'audio':
            # (B, T, F) fbank -> (B, 1, F, T)   (Audio/src/models/ast_models.py:152-153)
            data = data.unsqueeze(1).transpose(2, 3)
        if self.modality == 'graph':
            return self.embed(data)[0]
        return self.embed(data)
