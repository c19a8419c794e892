"""TEST INFRASTRUCTURE: restatement of the two consumers of the report text, so that the format the
drop-in CLI prints can be checked against what they expect without /root/reference at run time.

They read the all-lanes summary (``tail`` of the per-lane reports, Snakefile.count_and_push:166-172) line by
line and emit Confluence wiki markup (summary_to_wiki.py:16-44) or an HTML table (summary_to_wiki2.py:20-76).
tests/golden/wiki/ holds what the unmodified scripts printed for the same input (tests/golden/make_golden.py);
test_workflow.py pins this restatement to those files."""
import re


def _fmtline(line):
    """summary_to_wiki.py:35-44: the values of a Level line as a table row, the last box of level 1 in red."""
    level = re.match(r"Level: (\d+)", line).group(1)
    vals = [item.split(": ", 1)[1] for item in line.split("\t")]
    if level == "1":
        vals[-1] = "{color:red}%s{color}" % vals[-1]
    return "|".join([""] + vals + [""])


def to_wiki(text):
    """summary_to_wiki.py:16-33 over every line of ``text``."""
    out = []
    for line in text.split("\n")[:-1] if text.endswith("\n") else text.split("\n"):
        m = re.search(r"==> [\w/]*?(\d+)targets_lane(\d+)([TB]?)\.txt <==", line)
        if m:
            out.append("h3. %s targets per tile on lane %s%s" % m.groups())
            continue
        m = re.match(r"LaneSummary:.*(Tiles:.*)", line)
        if m:
            out.append(re.sub(r"\t", "   ", m.group(1)))
            continue
        if re.match(r"Level: 1\s", line):
            headings = [item.split(": ")[0] for item in line.split("\t")]
            out.append("||".join([""] + headings + [""]) + "\n" + _fmtline(line))
            continue
        if re.match("Level:", line):
            out.append(_fmtline(line))
    return "".join(x + "\n" for x in out)


def to_wiki2(text):
    """summary_to_wiki2.py:20-76: one table row per lane -- overall duplication, Picard-equivalent v1."""
    rows = ["<h3>Well Duplicates Summary</h3>",
            "<table>\n<tr>" + "".join("<th>%s</th>" % h for h in ("Lane", "Est. Duplication", "P.E. Scaled")) + "</tr>"]
    lane, raw = "0", "-"
    for line in text.split("\n")[:-1] if text.endswith("\n") else text.split("\n"):
        m = re.search(r"(\d+[TB]?)\.txt <==$", line)
        if m:
            lane = m.group(1)
        m = re.match(r"Overall duplication .*: ([0-9.%]+)", line)
        if m:
            raw = m.group(1).replace("%", " %")
        m = re.match(r"Picard-equivalent duplication v1: *([0-9.%]+)", line)
        if m:
            rows.append("<tr>" + "".join("<td>%s</td>" % a for a in (lane, raw, m.group(1).replace("%", " %"))) + "</tr>")
            raw = "-"
    rows.append("</table>")
    return "".join(x + "\n" for x in rows)
