"""Drop-in for the reference's ``gridsearch.py`` (lines 1-31): the 3 x 3 x 3 sweep over the reward-net regulariser
and the two fully connected widths, each point a full ``AC_IRL.outerloop()`` followed by ``test_reward_network()``,
one CSV line per point.  Run as a script from a directory holding the reference's data folders, or call ``run`` with
in-memory data (``ac_kwargs``) and shortened loops (``outerloop_kwargs``).
"""
from __future__ import annotations

import os

from . import ac_irl

LIST_REG = ['dropout', 'l1l2', 'dropout_l1l2']            # gridsearch.py:8
LIST_NFC3 = range(4, 10, 2)                                 # gridsearch.py:9
LIST_NFC4 = range(4, 10, 2)                                 # gridsearch.py:10


def run(outfile='results/reward_gridsearch_10_22.csv', list_reg=LIST_REG, list_nfc3=LIST_NFC3, list_nfc4=LIST_NFC4,
        theta=6.5, ac_kwargs=None, outerloop_kwargs=None):
    """Returns the rows written: (reg, n_fc3, n_fc4, reward_demo_avg_train, reward_demo_avg_test, reward_gen_avg, theta)."""
    ac_kwargs = dict(ac_kwargs or {})
    outerloop_kwargs = dict(outerloop_kwargs or {})
    os.makedirs(os.path.dirname(outfile) or ".", exist_ok=True)
    with open(outfile, 'w') as f:
        f.write('reg,n_fc3,n_fc4,reward_demo_avg_train,reward_demo_avg_test,reward_gen_avg,theta\n')
    rows = []
    for reg in list_reg:
        for n_fc3 in list_nfc3:
            for n_fc4 in list_nfc4:
                print("---------- reg = %s | n_fc3 = %d | n_fc4 = %d ----------" % (reg, n_fc3, n_fc4))
                ac = ac_irl.AC_IRL(theta=theta, reg=reg, n_fc3=n_fc3, n_fc4=n_fc4, **ac_kwargs)
                final_theta = ac.outerloop(**outerloop_kwargs)
                train, test, gen = ac.test_reward_network()
                with open(outfile, 'a') as f:
                    f.write('%s,%d,%d,%f,%f,%f,%f\n' % (reg, n_fc3, n_fc4, train, test, gen, final_theta))
                rows.append((reg, n_fc3, n_fc4, train, test, gen, final_theta))
    return rows


if __name__ == "__main__":
    run()
