"""Composable HTML renderers for `simple_wer_v2` (ref
`lingvo/tasks/asr/tools/custom_html_handlers.py`)."""

from __future__ import annotations

from lingvo_b200.models.asr.tools import simple_wer_v2


class ChainOfHtmlHandlers(simple_wer_v2.HtmlHandler):
  """Feeds each aligned pair through a list of handlers; the first that renders
  (returns a non-empty string) wins, the others still observe the pair (ref :22)."""

  def __init__(self, *handlers):
    super().__init__()
    self._handlers = list(handlers)

  def Setup(self, hypothesis, reference):
    for h in self._handlers:
      h.Setup(hypothesis, reference)

  def Render(self, hyp_word, ref_word, err_type):
    out = ''
    for h in self._handlers:
      piece = h.Render(hyp_word, ref_word, err_type)
      if piece and not out:
        out = piece
    return out


class TagHtmlHandler(simple_wer_v2.HtmlHandler):
  """Wraps words carrying one of `tags` (e.g. '<b>') in a styled span (ref :56)."""

  def __init__(self, tags=('<unk>',), color='lightgray'):
    super().__init__()
    self._tags = set(tags)
    self._color = color

  def Render(self, hyp_word, ref_word, err_type):
    word = hyp_word or ref_word
    if word in self._tags:
      return '<span style="background-color: %s">%s</span> ' % (self._color, word)
    return ''


class NewlineHtmlHandler(simple_wer_v2.HtmlHandler):
  """Turns a newline marker token into `<br>` (ref :93)."""

  def __init__(self, marker='<eol>'):
    super().__init__()
    self._marker = marker

  def Render(self, hyp_word, ref_word, err_type):
    if hyp_word == self._marker or ref_word == self._marker:
      return '<br>\n'
    return ''
