from pickle import *  # py2 cPickle stand-in (golden harness only)
