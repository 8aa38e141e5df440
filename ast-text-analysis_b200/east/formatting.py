# -*- coding: utf-8 -*-
"""Table / graph serialisation (mirror of east/formatting.py:4-80; format_table's NameError fixed)."""


def format_table(table, format):
    if format == "xml":
        return table2xml(table)
    elif format == "csv":
        return table2csv(table)
    raise Exception("Unknown table format: '%s'. Please use one of: 'xml', 'csv'." % format)


def table2xml(keyphrases_table):
    lines = ["<table>"]
    for keyphrase in sorted(keyphrases_table.keys()):
        lines.append('  <keyphrase value="%s">' % keyphrase)
        for text in sorted(keyphrases_table[keyphrase].keys()):
            lines.append('    <text name="%s">%.3f</text>' % (text, keyphrases_table[keyphrase][text]))
        lines.append("  </keyphrase>")
    lines.append("</table>")
    return "\n".join(lines) + "\n"


def table2csv(keyphrases_table):
    def quote(s):
        return '"' + s.replace('"', "'") + '"'

    keyphrases = sorted(keyphrases_table.keys())
    texts = sorted(keyphrases_table[keyphrases[0]].keys())
    rows = ["," + ",".join(quote(k) for k in keyphrases)]
    for text in texts:
        rows.append(quote(text) + "," + ",".join("%.3f" % keyphrases_table[k][text] for k in keyphrases))
    return "\n".join(rows) + "\n"


def format_graph(graph, format):
    if format == "gml":
        return graph2gml(graph)
    elif format == "edges":
        return graph2edges(graph)
    raise Exception("Unknown graph format: '%s'. Please use one of: 'gml', 'edges'." % format)


def graph2edges(graph):
    label_of = {node["id"]: node["label"] for node in graph["nodes"]}
    targets = {}
    for edge in graph["edges"]:
        targets.setdefault(label_of[edge["source"]], []).append(label_of[edge["target"]])
    return "".join("%s -> %s\n" % (src, ", ".join(dst)) for src, dst in targets.items())


def graph2gml(graph):
    out = ["graph", "[", "  directed 1",
           "  referral_confidence %.2f" % graph["referral_confidence"],
           "  relevance_threshold %.2f" % graph["relevance_threshold"],
           "  support_threshold %i" % graph["support_threshold"]]
    for node in graph["nodes"]:
        out += ["  node", "  [", "    id %i" % node["id"], '    label "%s"' % node["label"], "  ]"]
    for edge in graph["edges"]:
        out += ["  edge", "  [", "    source %i" % edge["source"], "    target %i" % edge["target"],
                "    confidence %.2f" % edge["confidence"], "  ]"]
    out.append("]")
    return "\n".join(out) + "\n"
