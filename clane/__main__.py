"""Alias of clane_b200.__main__ (drop-in for ``python -m clane``)."""
from clane_b200.__main__ import embedding, get_parser, main  # noqa: F401

if __name__ == "__main__":
    main()
