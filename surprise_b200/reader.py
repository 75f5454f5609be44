"""Reader: line format + rating scale (reference: surprise/reader.py:13-104)."""

_BUILTIN = {
    "ml-100k": dict(line_format="user item rating timestamp", rating_scale=(1, 5), sep="\t"),
    "ml-1m": dict(line_format="user item rating timestamp", rating_scale=(1, 5), sep="::"),
    "ml-20m": dict(line_format="user item rating timestamp", rating_scale=(0.5, 5.0), sep=",", skip_lines=1),
    "jester": dict(line_format="user item rating", rating_scale=(-10, 10)),
}


class Reader(object):
    """Parses rating lines.  Ratings are shifted by ``offset`` so that they are all >= 1 when the
    scale's lower bound is <= 0 (reader.py:58-59), as the reference does."""

    def __init__(self, name=None, line_format="user item rating", sep=None, rating_scale=(1, 5), skip_lines=0):
        if name:
            if name not in _BUILTIN:
                raise ValueError("unknown reader " + str(name) + ". Accepted values are " +
                                 ", ".join(_BUILTIN.keys()) + ".")
            self.__init__(**_BUILTIN[name])
            return
        self.sep = sep
        self.skip_lines = skip_lines
        self.rating_scale = rating_scale
        low = rating_scale[0]
        self.offset = 1 - low if low <= 0 else 0
        fields = line_format.split()
        wanted = ["user", "item", "rating"]
        self.with_timestamp = "timestamp" in fields
        if self.with_timestamp:
            wanted.append("timestamp")
        if any(f not in wanted for f in fields):
            raise ValueError("line_format parameter is incorrect.")
        self.indexes = [fields.index(w) for w in wanted]

    def parse_line(self, line):
        parts = line.split(self.sep)
        try:
            vals = [parts[k].strip() for k in self.indexes]
        except IndexError:
            raise ValueError("Impossible to parse line. Check the line_format and sep parameters.")
        ts = vals[3] if self.with_timestamp else None
        return vals[0], vals[1], float(vals[2]) + self.offset, ts
