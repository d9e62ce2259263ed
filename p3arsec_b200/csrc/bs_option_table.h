/*
 * bs_option_table.h -- the 1000-row base table of the synthetic option generator.
 *
 * PARSEC's inputgen builds every blackscholes input (in_4K.txt ... in_10M.txt) by repeating a fixed
 * 1000-row table cyclically: row i of the file is table[i % 1000].  That table (optionData.txt) is
 * PARSEC-owned and absent from the P3ARSEC overlay, so p3arsec_b200/data/optionData.txt is a stand-in
 * in the same initialiser syntax, drawn once with a fixed seed (tests/golden/make_option_table.py,
 * distribution in SURVEY.md 8d).  The field order is the reference's OptionData
 * (parsec-ff/pkgs/apps/blackscholes/src/blackscholes.c:89-100).
 */
#ifndef BS_OPTION_TABLE_H
#define BS_OPTION_TABLE_H

#define BS_TABLE_ROWS 1000

typedef struct bs_table_row {
    double s;        /* spot price                      */
    double strike;   /* strike price                    */
    double r;        /* risk-free rate                  */
    double divq;     /* dividend rate (unused)          */
    double v;        /* volatility                      */
    double t;        /* time to maturity, years         */
    char option_type; /* 'P' = put, anything else call  */
    double divs;     /* dividend values (unused)        */
    double dgrefval; /* DerivaGem reference value       */
} bs_table_row;

static const bs_table_row bs_option_table[BS_TABLE_ROWS] = {
#include "../data/optionData.txt"
};

#endif
