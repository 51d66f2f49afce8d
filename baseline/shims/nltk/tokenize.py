import re


class RegexpTokenizer:   # the one class datasets.py uses (RegexpTokenizer(r'\w+').tokenize)
    def __init__(self, pattern):
        self.p = re.compile(pattern)

    def tokenize(self, s):
        return self.p.findall(s)
