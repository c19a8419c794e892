"""Stand-in for the third-party ``Levenshtein`` module (PyPI python-Levenshtein).

TEST INFRASTRUCTURE ONLY.  The reference imports ``Levenshtein`` at
count_well_duplicates.py:9 and calls ``.distance`` / ``.hamming`` at :200,:252;
the module is a C extension that is not vendored in the reference, has no
pinned version there and is not installable offline, so its published
semantics are restated here from the textbook definitions:

* ``distance(a, b)``: unit-cost insert / delete / substitute edit distance
  (Wagner-Fischer dynamic programme).
* ``hamming(a, b)``: number of positions at which two equal-length strings
  differ (the real module raises ValueError on unequal lengths).

``tests/golden/make_golden.py`` puts this directory on PYTHONPATH so the
unmodified reference scripts can run in the authoring container.
PARITY UNPINNED at this boundary: the reference holds no test vectors for it.
"""


def distance(a, b):
    if a == b:
        return 0
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def hamming(a, b):
    if len(a) != len(b):
        raise ValueError("hamming() needs strings of equal length")
    return sum(x != y for x, y in zip(a, b))
