"""Token-level similarity of every host file to the reference file of the same name (comments and docstrings removed).

Only meaningful where /root/reference exists (the development container); used to keep the host mirror an independent
implementation of the reference's API contract rather than a transliteration.  usage: python tools/token_similarity.py [min]"""

from __future__ import annotations

import difflib
import io
import pathlib
import sys
import tokenize

MINE = pathlib.Path(__file__).resolve().parents[1] / "refiners_b200"
THEIRS = pathlib.Path("/root/reference/src/refiners")


def tokens(path: pathlib.Path) -> list[str]:
    out: list[str] = []
    previous = tokenize.INDENT
    for tok in tokenize.generate_tokens(io.StringIO(path.read_text()).readline):
        if tok.type in (tokenize.COMMENT, tokenize.NL, tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT, tokenize.ENCODING):
            if tok.type in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT):
                previous = tok.type
            continue
        if tok.type == tokenize.STRING and previous in (tokenize.NEWLINE, tokenize.INDENT, tokenize.DEDENT):
            continue  # a docstring / bare string statement
        previous = tok.type
        out.append(tok.string)
    return out


def main() -> None:
    floor = float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
    rows = []
    for mine in MINE.rglob("*.py"):
        rel = mine.relative_to(MINE)
        theirs = THEIRS / rel
        if not theirs.exists() or mine.name == "__init__.py":
            continue
        a, b = tokens(mine), tokens(theirs)
        if min(len(a), len(b)) < 40:
            continue
        rows.append((difflib.SequenceMatcher(None, a, b, autojunk=False).ratio(), str(rel), len(a), len(b)))
    for ratio, rel, na, nb in sorted(rows, reverse=True):
        if ratio >= floor:
            print(f"{ratio:.2f}  {rel}  ({na} vs {nb} tokens)")


if __name__ == "__main__":
    main()
