"""pybedtools for the reference harness: the BedTool subset bin/ntjoin_assemble.py uses (:339, :634-641, :669-677).
Intervals are half-open (chrom, start, end) like BED; overlaps need at least one shared base like bedtools."""
__version__ = "0.9.1"


class Interval:
    def __init__(self, chrom, start, end, count=None):
        self.chrom, self.start, self.end, self.count = chrom, int(start), int(end), count

    def __repr__(self):
        return f"Interval({self.chrom}:{self.start}-{self.end})"


class BedTool:
    def __init__(self, source="", from_string=False):
        if isinstance(source, list):
            self.intervals = source
        elif from_string:
            self.intervals = [Interval(*line.split("\t")[:3]) for line in source.split("\n") if line.strip()]
        else:
            with open(source) as fh:
                self.intervals = [Interval(*line.rstrip("\n").split("\t")[:3]) for line in fh if line.strip()]

    def __iter__(self):
        return iter(self.intervals)

    def __len__(self):
        return len(self.intervals)

    def sort(self, genome=None, **_kw):
        if genome is not None:                              # chromosomes in genome order, then by start
            order = {c: i for i, c in enumerate(genome)}
            key = lambda iv: (order.get(iv.chrom, len(order)), iv.start, iv.end)     # noqa: E731
        else:
            key = lambda iv: (iv.chrom, iv.start, iv.end)                            # noqa: E731
        return BedTool(sorted(self.intervals, key=key))

    def intersect(self, b=None, c=False, wa=False, **_kw):
        """-c -wa: every interval of self with the number of intervals of b it overlaps"""
        out = []
        for iv in self.intervals:
            n = sum(1 for o in b.intervals if o.chrom == iv.chrom and o.start < iv.end and iv.start < o.end)
            out.append(Interval(iv.chrom, iv.start, iv.end, count=n))
        return BedTool(out)

    def complement(self, i=None, g=None, **_kw):
        """bedtools complement -i <i> -g <g>: the parts of every chromosome of g not covered by i, in g's order"""
        src = i if i is not None else self
        out = []
        for chrom, span in g.items():
            length = span[1] if isinstance(span, (tuple, list)) else int(span)
            at = 0
            for iv in sorted((x for x in src.intervals if x.chrom == chrom), key=lambda x: (x.start, x.end)):
                if iv.start > at:
                    out.append(Interval(chrom, at, min(iv.start, length)))
                at = max(at, iv.end)
            if at < length:
                out.append(Interval(chrom, at, length))
        return BedTool(out)

    def saveas(self, path, **_kw):
        with open(path, "w") as fh:
            for iv in self.intervals:
                fh.write(f"{iv.chrom}\t{iv.start}\t{iv.end}\n")
        return self
