// FM_B200.java -- carskit.alg.cars.adaptation.dependent.FM (FM.java) on the B200 engine.
// FM.java keeps w0, w, V, p, k, regLw, regLf PRIVATE, so a subclass of FM could not hand them to the engine or read
// them back: this class extends ContextRecommender like FM does and restates the three small members around the hot
// loop -- the constructor (FM.java:49-55), initModel() (:57-74) and predict() (:93-113) -- while buildModel() (:115-220,
// the O(k * p * size) ALS sweep through a Guava table) runs on the device.
//     case "fm_b200": return new FM_B200(trainMatrix, testMatrix, fold);      setting.conf: FM_B200=-lw 0.01 -lf 0.02
package carskit.alg.b200;

import carskit.b200.B200;
import carskit.b200.Native;
import carskit.data.structure.DenseMatrix;
import carskit.data.structure.DenseVector;
import carskit.data.structure.SparseMatrix;
import carskit.generic.ContextRecommender;

public class FM_B200 extends ContextRecommender {
    private double w0;
    private int p, k;
    private DenseVector w; // [p]
    private DenseMatrix V; // [p x k]
    private final float regLw, regLf;

    public FM_B200(SparseMatrix trainMatrix, SparseMatrix testMatrix, int fold) {
        super(trainMatrix, testMatrix, fold);
        this.algoName = "FM_B200";
        regLw = algoOptions.getFloat("-lw");
        regLf = algoOptions.getFloat("-lf");
    }

    /** FM.initModel (FM.java:57-74): w0 = 0, w ~ U(0, 1), V ~ N(initMean, initStd). The row cache Q lives on the device. */
    @Override
    protected void initModel() throws Exception {
        super.initModel();
        k = numFactors;
        p = numUsers + numItems + numConditions;
        w0 = 0.0;
        w = new DenseVector(p);
        w.init();
        V = new DenseMatrix(p, k);
        V.init(initMean, initStd);
    }

    /**
     * FM.predict (FM.java:93-113) restricted to the three non-zero features of a rating (x_u = 1, x_{U+j} = 1,
     * x_{U+I+ctx} = 1 / numContextDims while that index is < p, FM.java:76-91). The reference adds p terms of which all
     * others are exactly zero, in ascending index order -- the same additions in the same order.
     */
    @Override
    protected double predict(int u, int j, int c) throws Exception {
        int[] idx = { u, numUsers + j, numUsers + numItems + c }; // ascending, like the reference's `for (j = 0; j < p; ++j)`
        double[] x = { 1.0, 1.0, 1.0 / rateDao.numContextDims() };
        int n = idx[2] < p ? 3 : 2; // FM.java:81-87: the context feature only exists while its index is < p
        double pred = w0;
        for (int a = 0; a < n; a++)
            pred += w.get(idx[a]) * x[a];
        double sum = 0.0;
        for (int f = 0; f < k; f++) {
            double sum1 = 0.0, sum2 = 0.0;
            for (int a = 0; a < n; a++) {
                double dot = V.get(idx[a], f) * x[a];
                sum1 += dot;
                sum2 += Math.pow(dot, 2);
            }
            sum += Math.pow(sum1, 2) - sum2;
        }
        pred += 0.5 * sum;
        return pred;
    }

    /** Replaces FM.buildModel() (FM.java:115-220): pre-pass + numIters ALS iterations; no isConverged(), no learn.rate. */
    @Override
    protected void buildModel() throws Exception {
        B200.Ratings x = B200.flattenContextual(trainMatrix, rateDao);
        double[] fw = B200.flatten(w), fV = B200.flatten(V);
        long h = Native.fmCreate(numUsers, numItems, numConditions, rateDao.numContexts(), k, rateDao.numContextDims(),
                x.u, x.j, x.ctx, x.r, (double) regLw, (double) regLf, B200.devicesFor(fold, 1)[0]);
        try {
            Native.fmUploadAndPrepare(h, w0, fw, fV); // errors[n] = r - predict, Q[n][f] (FM.java:118-146)
            for (int iter = 1; iter <= numIters; iter++)
                loss = Native.fmIteration(h);         // w0 step, w steps, V steps (FM.java:148-219)
            w0 = Native.fmDownload(h, fw, fV);
        } finally {
            Native.fmDestroy(h);
        }
        B200.unflatten(fw, w);
        B200.unflatten(fV, V);
    }
}
