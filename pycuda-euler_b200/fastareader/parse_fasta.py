"""fastareader.parse_fasta -- FASTA record iterator with the reference's interface
(src/fastareader/parse_fasta.py: ``Dna`` records, ``Fasta(handle)`` iterator, multi-line
sequences joined).  Pinned by the reference's tests/test_fasta_reader.py."""


class Dna:
    """One FASTA record: ``head`` (text after '>') and ``seq`` (list of sequence lines)."""

    def __init__(self, header, sequence):
        self.head = header
        self.seq = sequence

    def __repr__(self):
        return '<Dna %s>' % self.head

    def __str__(self, separator=''):
        return '>%s\n%s' % (self.head, separator.join(self.seq))

    def __len__(self):
        return sum(len(s) for s in self.seq)

    @property
    def sequence(self):
        return ''.join(self.seq)


class Fasta:
    """Iterate over the records of an open FASTA handle."""

    def __init__(self, handle):
        self.handle = handle

    def __repr__(self):
        return '<Fasta %r>' % (self.handle,)

    def __iter__(self):
        header, lines, seen = '', [], False
        for line in self.handle:
            if line.startswith('>'):
                if lines:
                    yield Dna(header, lines)
                header, lines, seen = line[1:].rstrip('\n'), [], True
            else:
                lines.append(line.strip())
        if seen or lines:
            yield Dna(header, lines)
