package gammaspike.logger;

import java.io.PrintStream;

import gammaspike.b200.GsmB200;
import gammaspike.b200.GsmContext;

/** ProportionOfSaltation from the sums the evaluation of the state already produced (UNCOMPILED here).  The column header
 *  (init, SaltativeProportionLogger.java:27-30) is inherited. */
public class SaltativeProportionLoggerB200 extends SaltativeProportionLogger {
    /** SPL:33-61: abruptChange / totalChange over the branches with len > 0 that are neither sampled ancestors nor the root */
    @Override
    public void log(long sample, PrintStream out) {
        final double[] row = GsmContext.of(treeInput.get()).eval();
        out.print(row[GsmB200.O_SUM_BURST] / row[GsmB200.O_SUM_DIST] + "\t");
    }
}
