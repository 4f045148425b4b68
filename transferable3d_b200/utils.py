"""sunrgbd/sunrgbd_data/utils.py:341-348: the gz-pickle container of the prepared frustum files
(sunrgbd_data.py:193-195 writes the 13-list training layout, :325-326 the 7-list rgb-detection layout).
The files were written by Python 2 cPickle; encoding='latin1' reads their numpy arrays and str objects under Python 3."""
import gzip
import pickle


def save_zipped_pickle(obj, filename, protocol=2):
    with gzip.open(filename, 'wb') as f:
        pickle.dump(obj, f, protocol)


def load_zipped_pickle(filename):
    with gzip.open(filename, 'rb') as f:
        return pickle.load(f, encoding='latin1')
