// CAMF_ICS_B200.java -- CAMF_ICS (src/carskit/alg/cars/adaptation/dependent/sim/CAMF_ICS.java) with buildModel() on the
// B200 engine (EXACT mode: one chain through ccMatrix_ICS, so one warp -- meant for the small data sets the model targets).
//     case "camf_ics_b200": return new CAMF_ICS_B200(trainMatrix, testMatrix, fold);
package carskit.alg.b200;

import java.util.ArrayList;
import java.util.List;

import carskit.alg.cars.adaptation.dependent.sim.CAMF_ICS;
import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.SparseMatrix;

public class CAMF_ICS_B200 extends CAMF_ICS {
    public CAMF_ICS_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "CAMF_ICS_B200";
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

    /** Replaces the per-rating loop of CAMF_ICS.buildModel() (CAMF_ICS.java:60-129). */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        List<List<Integer>> conds = new ArrayList<>();
        for (int c = 0; c < rateDao.numContexts(); c++)
            conds.add(getConditions(c));
        int[][] ctx = B200.contextTable(conds);
        int C = numConditions;
        double[] fP = B200.flatten(P), fQ = B200.flatten(Q), cc = new double[C * C];
        for (int a = 0; a < C; a++)
            for (int b = 0; b < C; b++)
                cc[a * C + b] = ccMatrix_ICS.get(a, b); // librec SymmMatrix: symmetric read
        int[] empty = new int[EmptyContextConditions.size()];
        for (int i = 0; i < empty.length; i++)
            empty[i] = EmptyContextConditions.get(i);
        B200.train(Native.CAMF_ICS, Native.EXACT, numUsers, numItems, C, numFactors, x, ctx, globalMean,
                (double) regU, (double) regI, (double) regB, (double) regC, B200.devicesFor(fold, 1), numIters, control,
                fP, fQ, null, null, null, null, null, cc, empty);
        B200.unflatten(fP, P);
        B200.unflatten(fQ, Q);
        for (int a = 0; a < C; a++)
            for (int b = 0; b <= a; b++)
                ccMatrix_ICS.set(a, b, cc[a * C + b]);
    }
}
