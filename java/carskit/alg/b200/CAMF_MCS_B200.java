// CAMF_MCS_B200.java -- CAMF_MCS (src/carskit/alg/cars/adaptation/dependent/sim/CAMF_MCS.java) with buildModel() on the
// B200 engine (EXACT mode: one chain through cVector_MCS, so one warp).  The engine derives upbound = 1 / sqrt(numContextDims)
// and lowbound = 1 / 10^100 (CAMF_MCS.java:44-45, private there) from the number of "na" conditions, and scales the
// loss by 0.05 like the reference (:158).
//     case "camf_mcs_b200": return new CAMF_MCS_B200(trainMatrix, testMatrix, fold);
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.cars.adaptation.dependent.sim.CAMF_MCS;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class CAMF_MCS_B200 extends CAMF_MCS {
    public CAMF_MCS_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "CAMF_MCS_B200";
    }

    private final B200.EpochControl control = new B200.EpochControl() {
        public double lRate() {
            return lRate;
        }

        public boolean afterEpoch(int iter, double epochLoss) throws Exception {
            loss = epochLoss;
            return isConverged(iter);
        }
    };

    /** Replaces the per-rating loop of CAMF_MCS.buildModel() (CAMF_MCS.java:71-167). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        List<List<Integer>> conds = new ArrayList<>();
        for (int c = 0; c < rateDao.numContexts(); c++)
            conds.add(getConditions(c));
        int[][] ctx = B200.contextTable(conds);
        double[] fP = B200.flatten(P), fQ = B200.flatten(Q), pos = cVector_MCS.getData().clone();
        int[] empty = new int[EmptyContextConditions.size()];
        for (int i = 0; i < empty.length; i++)
            empty[i] = EmptyContextConditions.get(i);
        B200.train(Native.CAMF_MCS, Native.EXACT, numUsers, numItems, numConditions, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, B200.devicesFor(fold, 1), numIters, control,
                fP, fQ, null, null, null, null, null, pos, empty, 0);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        System.arraycopy(pos, 0, cVector_MCS.getData(), 0, pos.length);
    }
}
