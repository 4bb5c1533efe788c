"""Byte offsets and segment ids of the tokenizers (reference tokenizers.py:95, :134)."""

import torch

from lingvo_b200.core import tokenizers


def _Ascii(**kw):
  return tokenizers.AsciiTokenizer.Params().Set(name='tok', **kw).Instantiate()


def test_ascii_offsets_are_byte_positions():
  tok = _Ascii(append_eos=True)
  ids, labels, paddings, start, end = tok.StringsToIdsWithOffsets(['hi you', 'a'], 8)
  assert ids.shape == labels.shape == paddings.shape == start.shape == end.shape
  n0 = int((1 - paddings[0]).sum())
  assert n0 == 7                                            # 6 chars + eos
  assert start[0, :6].tolist() == [0, 1, 2, 3, 4, 5]
  assert end[0, :6].tolist() == [1, 2, 3, 4, 5, 6]
  assert start[0, 6] == end[0, 6] == 6                      # eos: empty span at the end
  assert start[1, 0] == 0 and end[1, 0] == 1
  plain = tok.StringsToIds(['hi you', 'a'], 8)
  assert torch.equal(plain[0], ids) and torch.equal(plain[1], labels)


def test_vocab_tokenizer_offsets_skip_whitespace(tmp_path):
  vocab = tmp_path / 'vocab.txt'
  vocab.write_text('<unk>\n<s>\n</s>\nhello\nworld\n')
  tok = tokenizers.VocabFileTokenizer.Params().Set(
      name='tok', token_vocab_filepath=str(vocab), append_eos=False).Instantiate()
  _, labels, paddings, start, end = tok.StringsToIdsWithOffsets(['hello   world'], 4)
  assert int((1 - paddings[0]).sum()) == 2
  assert start[0, :2].tolist() == [0, 8] and end[0, :2].tolist() == [5, 13]
  assert labels[0, :2].tolist() == [3, 4]


def test_segments_and_external_append_eos():
  tok = _Ascii(append_eos=False)
  ids, labels, paddings, seg = tok.StringsToIdsWithSegments(['ab<segment>c<segment>de'], 8,
                                                            external_append_eos=True)
  n = int((1 - paddings[0]).sum())
  assert n == 6                                              # a b c d e eos
  assert seg[0, :n].tolist() == [0, 0, 1, 2, 2, 2]
  assert labels[0, n - 1] == tok.params.target_eos_id
  _, _, pad_no_eos, _ = tok.StringsToIdsWithSegments(['ab<segment>c'], 8)
  assert int((1 - pad_no_eos[0]).sum()) == 3
  tok.Initialize()
